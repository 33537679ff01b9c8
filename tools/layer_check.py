"""Kernel-experiment helper (not part of the product or the tests): runs the wide implicit-GEMM layers
of the model at full size (per-GPU batch B, decoders over 4 groups) on random fp16 operands and prints
 * a checksum of every output (3 repetitions: 'stable' = bit-identical from launch to launch), so two builds
   or two code paths can be compared bit for bit, e.g.
       python tools/layer_check.py 1024 > a.txt; MMDYN_IGEMM_PAIRS=1 python tools/layer_check.py 1024 > b.txt
 * summary statistics of the split-K (fp32 atomics) layers, and
 * CUDA-event timings of the same layers (TIMES_US line), e.g. against another build selected with
   MMDYN_B200_LIB=/path/to/libmmdyn_experiment.so."""
import hashlib, sys, torch
sys.path.insert(0, ".")
from mmdyn_b200 import engine, ops
from mmdyn_b200.pytorch.models.models import setup_model
torch.manual_seed(0)
KW = dict(condition_dim=0, input_dim=4096, architecture="cnn", conditional=False, categorical_conditions=False, latent_size=256)
m = setup_model("cnn-vae", **KW).cuda()
arena, ex = engine.get_execs(m, torch.device("cuda"))
enc, dec = ex["enc"]["encoder"], ex["dec"]["decoder"]
enc.refresh()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(device="cuda").manual_seed(1)
def rnd(*s): return (torch.rand(*s, device="cuda", generator=g) - 0.5).half()
def h(t): return hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()[:12]
def run(pl, which, a, out_shape, dtype=torch.float16, bias=None, reps=3):
    geom, W = (pl.lp.fwd, pl.Wf) if which == "fwd" else (pl.lp.dgrad, pl.Wd)
    hs = []
    for _ in range(reps):
        out = torch.zeros(out_shape, dtype=dtype, device="cuda")
        ops.igemm(geom, a, W, out, a.shape[0], bias=bias, out_mode=1 if dtype == torch.float32 and geom.out_mode == 0 else None)
        torch.cuda.synchronize()
        hs.append(h(out))
    return hs
res = {}
res["conv3.fwd"] = run(enc.c3, "fwd", rnd(B, 16, 16, 64), (B, 8, 8, 128))
res["conv4.fwd"] = run(enc.c4, "fwd", rnd(B, 8, 8, 128), (B, 5, 5, 256))
res["conv4.dgrad"] = run(enc.c4, "dgrad", rnd(B, 5, 5, 256), (B, 8, 8, 128))
res["conv3.dgrad"] = run(enc.c3, "dgrad", rnd(B, 8, 8, 128), (B, 16, 16, 64))
res["deconv1.fwd"] = run(dec.d1, "fwd", rnd(4 * B, 5, 5, 256), (4 * B, 8, 8, 128))
res["deconv1.dgrad"] = run(dec.d1, "dgrad", rnd(4 * B, 8, 8, 128), (4 * B, 5, 5, 256))
res["deconv2.fwd"] = run(dec.d2, "fwd", rnd(4 * B, 8, 8, 128), (4 * B, 16, 16, 64))
res["deconv2.dgrad"] = run(dec.d2, "dgrad", rnd(4 * B, 16, 16, 64), (4 * B, 8, 8, 128))
res["deconv3.fwd"] = run(dec.d3, "fwd", rnd(4 * B, 16, 16, 64), (4 * B, 32, 32, 32))
res["up.fwd"] = run(dec.up, "fwd", rnd(4 * B, 256), (4 * B, 5, 5, 256), bias=dec.up.bias)
res["heads.fwd"] = run(enc.heads, "fwd", rnd(4 * B, 512), (4 * B, 512), dtype=torch.float32, bias=enc.heads.bias)
for k, v in res.items():
    print(k, v[0], "stable" if len(set(v)) == 1 else "UNSTABLE " + " ".join(v))
# split-K fp32 outputs (atomics: compare numerically, not bitwise)
from mmdyn_b200 import plan
def run_ks(pl, which, a, out_shape, bias=None):
    geom, W = (pl.lp.fwd, pl.Wf) if which == "fwd" else (pl.lp.dgrad, pl.Wd)
    ks = plan.choose_ksplit(geom, a.shape[0])
    out = torch.zeros(out_shape, dtype=torch.float32, device="cuda")
    ops.igemm(geom, a, W, out, a.shape[0], bias=bias, ksplit=ks, out_mode=2 if ks > 1 else 1)
    torch.cuda.synchronize()
    return ks, out
for name, pl, which, a, shp, bias in (("fc.fwd", enc.fc, "fwd", rnd(B, 5, 5, 256), (B, 512), enc.fc.bias),
                                      ("up.dgrad", dec.up, "dgrad", rnd(4 * B, 5, 5, 256), (4 * B, 256), None),
                                      ("heads.dgrad", enc.heads, "dgrad", rnd(4 * B, 512), (4 * B, 512), None)):
    ks, o1 = run_ks(pl, which, a, shp, bias)
    geom, W = (pl.lp.fwd, pl.Wf) if which == "fwd" else (pl.lp.dgrad, pl.Wd)
    ref = a.reshape(a.shape[0], -1).float() @ W.float().t()[: a.reshape(a.shape[0], -1).shape[1]] if False else None
    print(name, "ksplit", ks, "sum %.6e abs %.6e" % (o1.double().sum().item(), o1.double().abs().sum().item()),
          "rows equal-ish:", float((o1[:16] - o1[-16:]).abs().max()))

# ---- timing of the same layers (CUDA events, 20 launches each) ----
def tm(pl, which, a, out_shape, dtype=torch.float16, bias=None):
    geom, W = (pl.lp.fwd, pl.Wf) if which == "fwd" else (pl.lp.dgrad, pl.Wd)
    out = torch.zeros(out_shape, dtype=dtype, device="cuda")
    f = lambda: ops.igemm(geom, a, W, out, a.shape[0], bias=bias, out_mode=1 if dtype == torch.float32 and geom.out_mode == 0 else None)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
T = {}
T["conv3.fwd"] = tm(enc.c3, "fwd", rnd(B, 16, 16, 64), (B, 8, 8, 128))
T["conv4.fwd"] = tm(enc.c4, "fwd", rnd(B, 8, 8, 128), (B, 5, 5, 256))
T["conv4.dgrad"] = tm(enc.c4, "dgrad", rnd(B, 5, 5, 256), (B, 8, 8, 128))
T["conv3.dgrad"] = tm(enc.c3, "dgrad", rnd(B, 8, 8, 128), (B, 16, 16, 64))
T["deconv1.fwd"] = tm(dec.d1, "fwd", rnd(4 * B, 5, 5, 256), (4 * B, 8, 8, 128))
T["deconv1.dgrad"] = tm(dec.d1, "dgrad", rnd(4 * B, 8, 8, 128), (4 * B, 5, 5, 256))
T["deconv2.fwd"] = tm(dec.d2, "fwd", rnd(4 * B, 8, 8, 128), (4 * B, 16, 16, 64))
T["deconv2.dgrad"] = tm(dec.d2, "dgrad", rnd(4 * B, 16, 16, 64), (4 * B, 8, 8, 128))
T["deconv3.fwd"] = tm(dec.d3, "fwd", rnd(4 * B, 16, 16, 64), (4 * B, 32, 32, 32))
T["deconv3.dgrad"] = tm(dec.d3, "dgrad", rnd(4 * B, 32, 32, 32), (4 * B, 16, 16, 64))
T["up.fwd"] = tm(dec.up, "fwd", rnd(4 * B, 256), (4 * B, 5, 5, 256), bias=dec.up.bias)
print("TIMES_US " + " ".join(f"{k}={v:.1f}" for k, v in T.items()))
