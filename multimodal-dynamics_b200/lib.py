"""ctypes binding of libmmdyn_b200.so (the C ABI declared in include/mmdyn_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is only used by callers for device memory and streams; this module passes raw pointers.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMDYN_B200_LIB: load another build of the same ABI (kernel experiments); default = the in-tree library
LIB_PATH = os.environ.get("MMDYN_B200_LIB") or os.path.join(_HERE, "libmmdyn_b200.so")
CSRC = os.path.join(_HERE, "csrc")

MAX_TAPS = 16
MAX_PHASES = 4
MAX_GROUPS = 8


class MmdynError(RuntimeError):
    pass


class IgemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("out", C.c_void_p), ("bias", C.c_void_p),
        ("n_img", C.c_int32), ("P", C.c_int32), ("OXv", C.c_int32),
        ("IH", C.c_int32), ("IW", C.c_int32), ("a_pix_stride", C.c_int32), ("Cin", C.c_int32),
        ("s_in", C.c_int32), ("ntaps", C.c_int32), ("n_phases", C.c_int32),
        ("tap_dy", (C.c_int8 * MAX_TAPS) * MAX_PHASES), ("tap_dx", (C.c_int8 * MAX_TAPS) * MAX_PHASES),
        ("N", C.c_int32), ("block_n", C.c_int32), ("ksplit", C.c_int32), ("row_mode", C.c_int32),
        ("out_mode", C.c_int32), ("OH", C.c_int32), ("OW", C.c_int32), ("s_out", C.c_int32),
        ("off_y", C.c_int32 * MAX_PHASES), ("off_x", C.c_int32 * MAX_PHASES), ("ldc", C.c_int32),
        ("a_row_stride", C.c_int32), ("a_img_stride", C.c_int32),
        ("bce_target", C.c_void_p), ("bce_mask", C.c_void_p), ("bce_dlogits", C.c_void_p), ("bce_loss", C.c_void_p),
        ("bce_gscale", C.c_float), ("bce_rows_per_group", C.c_int32), ("bce_slot", C.c_int32 * MAX_GROUPS),
        ("logit_row_lo", C.c_int32), ("logit_row_hi", C.c_int32),
        ("patch_mode", C.c_int32), ("bn_rows_per_group", C.c_int32), ("bn_sums", C.c_void_p),
        ("s_in_x", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("G", C.c_void_p), ("Nat", C.c_void_p), ("dW", C.c_void_p),
        ("n_img", C.c_int32), ("P", C.c_int32), ("OXv", C.c_int32), ("IH", C.c_int32), ("IW", C.c_int32),
        ("g_pix_stride", C.c_int32), ("Cg", C.c_int32), ("s_in", C.c_int32), ("ntaps", C.c_int32),
        ("tap_dy", C.c_int8 * MAX_TAPS), ("tap_dx", C.c_int8 * MAX_TAPS),
        ("Cn", C.c_int32), ("nat_stride", C.c_int32), ("ldw", C.c_int32), ("row_splits", C.c_int32),
        ("scale", C.c_float), ("g_row_stride", C.c_int32), ("g_img_stride", C.c_int32), ("s_in_x", C.c_int32),
    ]


class PoePass(C.Structure):
    _fields_ = [("mu_e", C.c_void_p * 4), ("lv_e", C.c_void_p * 4), ("n_experts", C.c_int32), ("eps", C.c_void_p),
                ("mu", C.c_void_p), ("lv", C.c_void_p), ("z", C.c_void_p), ("zh", C.c_void_p), ("zh2", C.c_void_p),
                ("kl_sum", C.c_void_p), ("dz", C.c_void_p * 3), ("dmu_in", C.c_void_p), ("dlv_in", C.c_void_p),
                ("dmu_e", C.c_void_p * 4), ("dlv_e", C.c_void_p * 4)]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_LL = C.c_longlong
_U64 = C.c_uint64

# name -> argtypes; every symbol declared in include/mmdyn_b200.h must appear here
SIGNATURES = {
    "mmdyn_last_error": ([], C.c_char_p),
    "mmdyn_version": ([], _I),
    "mmdyn_launch_count": ([], _LL),
    "mmdyn_init": ([_I], _I),
    "mmdyn_igemm": ([C.POINTER(IgemmDesc), _P], _I),
    "mmdyn_wgrad": ([C.POINTER(WgradDesc), _P], _I),
    "mmdyn_conv1_fwd": ([_P, _P, _P, _P, _I, _P], _I),
    "mmdyn_conv1_wgrad": ([_P, _P, _P, _I, _F, _I, _P], _I),
    "mmdyn_bn_stats": ([_P, _P, _I, _I, _I, _P], _I),
    "mmdyn_bn_finalize": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P, _P], _I),
    "mmdyn_bn_finalize_swish_fwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P], _I),
    "mmdyn_bn_swish_fwd": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "mmdyn_bn_swish_bwd_reduce": ([_P, _P, _P, _P, _P, _I, _I, _I, _P], _I),
    "mmdyn_bn_bwd_apply": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P], _I),
    "mmdyn_bn_bwd_apply_padded": ([_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _F, _P], _I),
    "mmdyn_swish_dropout_fwd": ([_P, C.POINTER(_P), _P, _I, _I, _I, _P], _I),
    "mmdyn_swish_dropout_bwd": ([_P, C.POINTER(_P), _P, _P, _I, _I, _I, _P], _I),
    "mmdyn_poe_fwd": ([C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P], _I),
    "mmdyn_poe_bwd": ([C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _P, C.POINTER(_P), _P, _P, _F, C.POINTER(_P), C.POINTER(_P),
                       _I, _I, _I, _I, _P], _I),
    "mmdyn_poe_fwd_multi": ([C.POINTER(PoePass), _I, _I, _I, _I, _I, _P], _I),
    "mmdyn_poe_bwd_multi": ([C.POINTER(PoePass), _I, _I, _I, _F, _I, _I, _I, _I, _P], _I),
    "mmdyn_bce_logits": ([_P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _P], _I),
    "mmdyn_mse": ([_P, _P, _P, _P, _F, _F, _I, _P], _I),
    "mmdyn_bce_logits_flat": ([_P, _P, _P, _P, _P, _P, _F, _I, _I, _P], _I),
    "mmdyn_mse_rows": ([_P, _P, _P, _F, _I, _I, _P], _I),
    "mmdyn_linear_f32_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "mmdyn_linear_f32_bwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P], _I),
    "mmdyn_colsum_f32": ([_P, _P, _I, _I, _I, _F, _P], _I),
    "mmdyn_colsum_f16": ([_P, _P, _I, _I, _I, _F, _P], _I),
    "mmdyn_pack_f16": ([_P, _P, _P, _LL, _P], _I),
    "mmdyn_split_f16": ([_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _F, _P], _I),
    "mmdyn_gather_f32": ([_P, _P, _P, _LL, _P], _I),
    "mmdyn_linear_f32_acc": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "mmdyn_linear_f32_wgrad": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P], _I),
    "mmdyn_cond_add_f16": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "mmdyn_cond_wgrad_f16": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], _I),
    "mmdyn_relu_f32": ([_P, _P, _LL, _P], _I),
    "mmdyn_act_grad_f32": ([_P, _P, _P, _I, _I, _I, _I, _P], _I),
    "mmdyn_resize_table_ints": ([_I, _I, _I, _I], _I),
    "mmdyn_resize_table": ([_I, _I, _I, _I, _P, _I], _I),
    "mmdyn_frames_u8_to_f32": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "mmdyn_unpack_add_f32": ([_P, _P, _P, _LL, _P], _I),
    "mmdyn_gather_add_f32": ([_P, _P, _P, _LL, _P], _I),
    "mmdyn_f32_to_f16": ([_P, _P, _LL, _F, _P], _I),
    "mmdyn_scale_f32": ([_P, _LL, _F, _P], _I),
    "mmdyn_logit_grad_pack": ([_P, _P, _F, _I, _I, _I, _I, _I, _P], _I),
    "mmdyn_adam_flat": ([_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _I, _F, _P], _I),
    "mmdyn_adam_flat_devstep": ([_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _P, _F, _P], _I),
    "mmdyn_sgd_flat": ([_P, _P, _P, _LL, _F, _F, _F, _I, _F, _P], _I),
    "mmdyn_adam_flat_guarded": ([_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _P, _F, _P, _P], _I),
    "mmdyn_sgd_flat_guarded": ([_P, _P, _P, _LL, _F, _F, _F, _I, _F, _P, _P], _I),
    "mmdyn_enable_peer_access": ([_I], _I),
    "mmdyn_ipc_export": ([_P, _P, C.POINTER(_LL)], _I),
    "mmdyn_ipc_import": ([_P, _LL, C.POINTER(_P)], _I),
    "mmdyn_peer_rs_adam_ag": ([C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, _P, _LL, _I, _I, _F, _F, _F, _F, _F, _P, _P,
                               _F, _P, _P, _P], _I),
    "mmdyn_fill_normal": ([_P, _LL, _U64, _U64, _P, _P], _I),
    "mmdyn_fill_dropout_mask": ([_P, _LL, _F, _U64, _U64, _P, _P], _I),
    "mmdyn_rng_advance": ([_P, _U64, _P], _I),
}

_lib = None
_inited_devices = set()


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libmmdyn_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise MmdynError("building libmmdyn_b200.so failed (see output above)")
    return LIB_PATH


def load():
    """Load the shared library and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmdynError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C multimodal-dynamics_b200/csrc`.  There is no CPU / PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error():
    return load().mmdyn_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise MmdynError(f"{what} failed (rc={rc}): {last_error()}")


def init(device_index):
    lib = load()
    if device_index not in _inited_devices:
        check(lib.mmdyn_init(int(device_index)), "mmdyn_init")
        _inited_devices.add(device_index)
    return lib


def launch_count():
    return int(load().mmdyn_launch_count())
