"""Execution engine of the cnn-vae / cnn-mvae step on top of the C-ABI kernels (ops.py).

Pieces
  ParamArena   one flat fp32 parameter arena + one flat gradient arena; every nn.Parameter of the
               model becomes a view (so torch optimizers / state_dict / .to() plumbing keep
               working, and the fused Adam and the gradient all-reduce see one buffer).
  PackedLayer  fp16 K-contiguous operand copies of the weights (forward and dgrad packing),
               refreshed by one gather kernel per layer whenever the arena changes.
  EncoderExec / DecoderExec / PoseExec
               forward + hand-written backward of the three network kinds of the reference
               (vae.py:179-301), batched over "groups" (= sub-sampled passes) with per-group
               BatchNorm statistics.
  StepEngine   the whole training / evaluation step of Reconstruction._evaluate_mvae and
               SeqModeling._evaluate_model (problems.py:473-546, 683-716): image-encoder trunks run
               ONCE and are shared by the passes that use them (only the Dropout mask differs),
               the loss-bearing decoder invocations of all passes run as one group-batched launch
               sequence, PoE/reparam/KL, losses, full backward into the gradient arena.

Gradients are carried through the fp16 backward tensors multiplied by `grad_scale` (default: the
batch size, which makes dlogits = sigmoid(x) - t exactly) and un-scaled in fp32 where they are
reduced into parameter gradients.
"""
import os

import numpy as np
import torch

from . import ops, plan

F16, F32 = torch.float16, torch.float32
LATENT_HEADS = 512  # [mu | logvar] of a 256-d latent


# ---------------------------------------------------------------------------------------------
# allocation helpers
# ---------------------------------------------------------------------------------------------
class FreshAlloc:
    """New tensors on every call (module-level API: activations are owned by the autograd ctx)."""

    def __init__(self, device):
        self.device = device

    def __call__(self, key, shape, dtype, zero=False):
        return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)


class Workspace:
    """Persistent named buffers (fused step: stable addresses, CUDA-graph friendly).

    A buffer is identified by (name, shape, dtype) and is NEVER replaced or freed once handed out: captured
    CUDA graphs hold raw pointers into these buffers, and several graphs (one per input shape, see
    Problem._graph_cache) as well as eager evaluations of other batch sizes share one Workspace.  A request
    for a known name with a new shape therefore allocates a second buffer next to the first instead of
    recycling it (the earlier graph keeps replaying into memory that is still its own).  `bufs[name]` is the
    buffer most recently handed out under that name.  `trim()` drops everything — only legal when no
    captured graph that used this workspace will be replayed again."""

    POOL_BYTES = 8 << 20

    def __init__(self, device):
        self.device = device
        self.bufs = {}       # name -> latest tensor
        self._all = {}       # (name, shape, dtype) -> tensor
        self.pools = []      # fp32 chunks holding every zero="step" buffer: ONE memset per chunk and step
        self.pool_used = 0   # floats used in the last chunk

    def __call__(self, key, shape, dtype, zero=False):
        """zero=True: zeroed on every call.  zero="step": fp32 accumulator that only needs to be zero when
        the step starts (BatchNorm sums, bias-gradient partials ...) — carved out of a pool that
        begin_step() clears with a single memset instead of one fill kernel per buffer."""
        shape = tuple(int(s) for s in shape)
        full = (key, shape, dtype)
        t = self._all.get(full)
        if t is None:
            if zero == "step" and dtype == F32:
                t = self._from_pool(shape)
            else:
                t = torch.zeros(shape, dtype=dtype, device=self.device)
            self._all[full] = t
        elif zero is True:
            t.zero_()
        self.bufs[key] = t
        return t

    def _from_pool(self, shape):
        n = 1
        for s in shape:
            n *= s
        n_al = (n + 63) // 64 * 64
        cap = self.POOL_BYTES // 4
        if n_al > cap:
            return torch.zeros(shape, dtype=F32, device=self.device)  # too big to pool; caller gets zero=True semantics
        if not self.pools or self.pool_used + n_al > cap:
            self.pools.append(torch.zeros(cap, dtype=F32, device=self.device))
            self.pool_used = 0
        t = self.pools[-1][self.pool_used:self.pool_used + n].view(shape)
        self.pool_used += n_al
        return t

    def begin_step(self):
        for p in self.pools:
            p.zero_()

    def trim(self):
        """Free every buffer (benchmark sweeps over many batch sizes).  The caller guarantees that no CUDA
        graph captured on this workspace is replayed afterwards."""
        self.bufs.clear()
        self._all.clear()
        self.pools = []
        self.pool_used = 0

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self._all.values())


# ---------------------------------------------------------------------------------------------
# parameter arena
# ---------------------------------------------------------------------------------------------
class ParamArena:
    ALIGN = 4  # floats (16 bytes)

    def __init__(self, module):
        self.module = module
        named = list(module.named_parameters())
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.offset, self.numel = {}, {}
        off = 0
        for n, p in named:
            self.offset[n] = off
            self.numel[n] = p.numel()
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.flat = None
        self.grad = None
        self.manual_version = 0

    def _is_bound(self, device):
        if self.flat is None or self.flat.device != device:
            return False
        base = self.flat.data_ptr()
        for n, p in zip(self.names, self.params):
            if p.data_ptr() != base + 4 * self.offset[n] or not p.is_contiguous():
                return False
        return True

    def ensure(self, device=None):
        """Make every parameter a view of the flat arena on `device` (idempotent)."""
        device = torch.device(device) if device is not None else self.params[0].device
        if device.type != "cuda":
            raise RuntimeError("mmdyn_b200 has no CPU path: move the model to a CUDA device (B200)")
        if self._is_bound(device):
            return False
        flat = torch.zeros(self.total, dtype=F32, device=device)
        with torch.no_grad():
            for n, p in zip(self.names, self.params):
                o, k = self.offset[n], self.numel[n]
                flat[o:o + k].copy_(p.data.reshape(-1))
                p.data = flat[o:o + k].view(p.shape)
        self.flat = flat
        self.grad = torch.zeros(self.total, dtype=F32, device=device)
        for p in self.params:
            p.grad = None
        self.manual_version += 1
        return True

    def view(self, name, of=None):
        o, k = self.offset[name], self.numel[name]
        return (self.flat if of is None else of)[o:o + k]

    def attach_grads(self):
        """p.grad := view of the gradient arena (zeroed where it was None, like a fresh backward)."""
        base = self.grad.data_ptr()
        missing = [i for i, p in enumerate(self.params)
                   if p.grad is None or p.grad.data_ptr() != base + 4 * self.offset[self.names[i]]]
        if not missing:
            return
        if len(missing) == len(self.params):
            self.grad.zero_()
        for i in missing:
            n, p = self.names[i], self.params[i]
            g = self.view(n, self.grad)
            if len(missing) != len(self.params):
                g.zero_()
            p.grad = g.view(p.shape)

    def token(self):
        return (id(self.flat), self.flat._version, self.manual_version)

    def bump(self):
        self.manual_version += 1


def get_arena(module, device=None):
    arena = module.__dict__.get("_mmdyn_arena")
    if arena is None:
        arena = ParamArena(module)
        module.__dict__["_mmdyn_arena"] = arena
    arena.ensure(device)
    return arena


# ---------------------------------------------------------------------------------------------
# packed layers
# ---------------------------------------------------------------------------------------------
class PackedLayer:
    def __init__(self, lp, device, cache):
        self.lp = lp
        self.device = device

        def up(a):
            if a is None:
                return None
            t = cache.get(id(a))
            if t is None:
                t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
                cache[id(a)] = t
            return t
        self.idx_fwd, self.idx_dgrad = up(lp.idx_fwd), up(lp.idx_dgrad)
        self.idx_wgrad, self.bias_idx = up(lp.idx_wgrad), up(lp.bias_idx)
        self.Wf = self.Wd = self.bias = None  # views into the ModelPacker arenas


class ModelPacker:
    """All fp16 operand copies (and permuted fp32 biases) of a model live in two arenas that are
    rebuilt by ONE gather kernel each when the parameter arena changes; layers hold views."""

    def __init__(self, nets, device):
        layers = [pl for net in nets for pl in net.layers]
        idx_parts, self.slots = [], []
        off = 0
        # Three contiguous stages in the order a training step needs them: forward copies of the encoder-side
        # networks (read by the step's first kernels), forward copies of the decoders, data-gradient copies.
        # refresh(staged=True) packs stage 0 on the calling stream and stages 1 / 2 on a side stream, so only
        # ~a quarter of this batch-size-independent gather sits on the step's critical path.
        first = [pl for net in nets if not isinstance(net, DecoderExec) for pl in net.layers]
        later = [pl for net in nets if isinstance(net, DecoderExec) for pl in net.layers]
        self.stage_off = [0]
        for group, attr in ((first, "Wf"), (later, "Wf"), (layers, "Wd")):
            for pl in group:
                idx = pl.idx_fwd if attr == "Wf" else pl.idx_dgrad
                if idx is not None:
                    # 16-byte aligned fp16 views (TMA operands, vector loads / stores of the pack kernel)
                    assert idx.numel() % 8 == 0, (pl.lp.name, attr, tuple(idx.shape))
                    idx_parts.append(idx.reshape(-1))
                    self.slots.append((pl, attr, off, tuple(idx.shape)))
                    off += idx.numel()
            self.stage_off.append(off)
        self._side, self._ev, self._pending = None, [None, None], [False, False]
        self.idx = torch.cat(idx_parts) if idx_parts else None
        self.W = torch.empty(off, dtype=F16, device=device)
        for pl, attr, o, shape in self.slots:
            setattr(pl, attr, self.W[o:o + shape[0] * shape[1]].view(shape))
        b_parts, boff = [], 0
        for pl in layers:
            if pl.bias_idx is not None:
                b_parts.append((pl, boff, pl.bias_idx.numel()))
                boff += pl.bias_idx.numel()
        self.bias_idx = torch.cat([pl.bias_idx for pl, _, _ in b_parts]) if b_parts else None
        self.bias = torch.empty(boff, dtype=F32, device=device)
        for pl, o, k in b_parts:
            pl.bias = self.bias[o:o + k]
        self.token = None

    def refresh(self, arena, force=False, staged=False):
        """staged: only the fused step asks for it, and then calls wait_stage(1) before its decoders and
        wait_stage(2) before it returns from the forward (so the side stream is always joined again)."""
        tok = arena.token()
        if force or tok != self.token:
            o = self.stage_off
            if self.idx is not None and staged and o[1] > 0 and o[3] > o[1]:
                cur = torch.cuda.current_stream()
                if self._side is None:
                    self._side = torch.cuda.Stream()
                ev = torch.cuda.Event()
                ev.record(cur)
                ops.pack_f16(arena.flat, self.idx[:o[1]], self.W[:o[1]])
                self._side.wait_event(ev)
                with torch.cuda.stream(self._side):
                    for s in (1, 2):
                        if o[s + 1] > o[s]:
                            ops.pack_f16(arena.flat, self.idx[o[s]:o[s + 1]], self.W[o[s]:o[s + 1]])
                        self._ev[s - 1] = torch.cuda.Event()
                        self._ev[s - 1].record(self._side)
                        self._pending[s - 1] = True
            elif self.idx is not None:
                ops.pack_f16(arena.flat, self.idx, self.W)
            if self.bias_idx is not None:
                ops.gather_f32(arena.flat, self.bias_idx, self.bias)
            self.token = tok

    def wait_stage(self, s):
        """The calling stream waits for stage s (1: decoder forward copies, 2: data-gradient copies) of a staged refresh."""
        for k in range(s):
            if self._pending[k]:
                torch.cuda.current_stream().wait_event(self._ev[k])
                self._pending[k] = False


class GradPack:
    """Packed fp32 weight-gradient segments of one sub-network: zeroed by one memset, filled by the
    wgrad launches, scattered into the gradient arena by ONE unpack kernel."""

    def __init__(self, net, device):
        self.seg, parts, off = {}, [], 0
        for pl in net.layers:
            if pl.lp.wgrad is not None:
                n = pl.lp.wgrad.Cn * pl.lp.wgrad.K
                self.seg[id(pl)] = (off, (pl.lp.wgrad.Cn, pl.lp.wgrad.K))
                parts.append(pl.idx_wgrad.reshape(-1))
                off += n
        for name, idx in getattr(net, "extra_wgrads", {}).items():
            self.seg[name] = (off, tuple(idx.shape))
            parts.append(idx.reshape(-1))
            off += idx.numel()
        self.idx = torch.cat(parts)
        self.buf = torch.zeros(off, dtype=F32, device=device)
        # inverse map over the arena range this sub-network's weights occupy: the flush then walks the
        # arena (coalesced read-modify-write) and gathers from the packed buffer instead of scattering
        idx_h = self.idx.cpu().numpy()
        pos = np.nonzero(idx_h >= 0)[0]
        tgt = idx_h[pos]
        self.lo, self.inv = 0, None
        if tgt.size and np.unique(tgt).size == tgt.size:
            self.lo, hi = int(tgt.min()), int(tgt.max()) + 1
            inv = np.full(hi - self.lo, -1, np.int32)
            inv[tgt - self.lo] = pos.astype(np.int32)
            self.inv = torch.from_numpy(inv).to(device)
        self.wstream = torch.cuda.Stream(device=device) if os.environ.get("MMDYN_SERIAL_BRANCHES") is None else None

    def begin(self):
        self.buf.zero_()

    def view(self, key):
        o, shape = self.seg[key]
        return self.buf[o:o + shape[0] * shape[1]].view(shape)

    def flush(self, arena):
        if self.wstream is not None:
            torch.cuda.current_stream().wait_stream(self.wstream)
        if self.inv is not None:
            ops.gather_add_f32(self.buf, self.inv, arena.grad[self.lo:self.lo + self.inv.numel()])
        else:
            ops.unpack_add_f32(self.buf, self.idx, arena.grad)


class _NetBase:
    def __init__(self, arena, prefix, device):
        self.arena, self.prefix, self.device = arena, prefix, device
        self.layers = []
        self._cache = {}
        self._token = None
        self.packer = None

    def _pl(self, lp):
        pl = PackedLayer(lp, self.device, self._cache)
        self.layers.append(pl)
        return pl

    def _full(self, name):
        return self.prefix + "." + name if self.prefix else name

    def off(self, name):
        return self.arena.offset[self._full(name)]

    def pview(self, name, grad=False):
        return self.arena.view(self._full(name), self.arena.grad if grad else None)

    def buf(self, name):
        return self.arena.module.get_buffer(self._full(name))

    def refresh(self):
        self.packer.refresh(self.arena)


def _ig(pl, which, A, out, n_img, bias=None, f32_out=False, bce=None, stats=None):
    """One implicit-GEMM launch of layer `pl` ('fwd' or 'dgrad'); fp32 outputs may split K.
    stats = (sums, images per group): BatchNorm statistics of the output in the epilogue (patch kernel)."""
    geom, W = (pl.lp.fwd, pl.Wf) if which == "fwd" else (pl.lp.dgrad, pl.Wd)
    kw = dict(tag=f"{pl.lp.name}.{which}", macs_per_img=pl.lp.extra.get("macs"))
    if bce is not None:
        kw["bce"] = bce
    if stats is not None:
        kw["stats"] = stats
    ks = plan.choose_ksplit(geom, n_img) if f32_out else 1
    if ks > 1:
        out.zero_()
        ops.igemm(geom, A, W, out, n_img, bias=bias, ksplit=ks, out_mode=2, **kw)
    else:
        ops.igemm(geom, A, W, out, n_img, bias=bias, out_mode=1 if f32_out else None, **kw)


class _on_side:
    """Launch the enclosed kernels on the GradPack's weight-gradient stream: a layer's wgrad only
    needs that layer's dRaw, so it overlaps the next layer's dgrad + BatchNorm backward."""

    def __init__(self, gp):
        self.st = getattr(gp, "wstream", None) if gp is not None else None

    def __enter__(self):
        if self.st is not None:
            self.st.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.st)
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.st is not None:
            self.ctx.__exit__(*a)


def _wgrad_into(pl, G, Nat, n_img, arena, alloc, key, scale, gp=None):
    """Weight gradient of layer `pl` into the arena: through the sub-network's GradPack (fused step:
    one memset + one scatter per sub-network) or a private scratch + scatter (module-level API)."""
    wg = pl.lp.wgrad
    dWp = gp.view(id(pl)) if gp is not None else alloc(key, (wg.Cn, wg.K), F32, zero=True)
    with _on_side(gp):
        ops.wgrad(wg, G, Nat, dWp, n_img, scale=scale, row_splits=plan.choose_row_splits(wg, n_img),
                  tag=f"{pl.lp.name}.wgrad", macs_per_img=pl.lp.extra.get("macs"))
    if gp is None:
        ops.unpack_add_f32(dWp, pl.idx_wgrad, arena.grad)


FUSE_STATS = os.environ.get("MMDYN_NO_FUSED_STATS") is None  # BatchNorm sums from the producing GEMM's epilogue
PAIR_GRADS = os.environ.get("MMDYN_NO_PAIR_GRADS") is None  # deconv3's gradient stored with an x border, read as pixel pairs
STAGED_PACK = os.environ.get("MMDYN_NO_STAGED_PACK") is None  # weight packing in three stages, two of them off the critical path


class _BN:
    """Grouped training-mode BatchNorm2d + Swish around a raw conv output."""

    def __init__(self, net, idx_name, C):
        self.net, self.name, self.C = net, idx_name, C

    def forward(self, raw, act, G, rows, alloc, key, track, repeat=1):
        st = self.alloc_fwd(G, alloc, key)
        self.run_fwd(raw, act, G, rows, st, 0, track, repeat)
        return st[1], st[2]

    def alloc_fwd(self, G, alloc, key):
        C = self.C
        return (alloc(key + ".sums", (G, C, 2), F32, zero="step"), alloc(key + ".ab", (G, C, 2), F32),
                alloc(key + ".mi", (G, C, 2), F32))

    def run_fwd(self, raw, act, Gc, rows, st, g0, track, repeat=1, have_stats=False):
        """Groups g0 .. g0+Gc-1: `raw` / `act` hold exactly these groups' rows, `st` = alloc_fwd(...) of all groups.
        have_stats: `sums` was already filled by the epilogue of the GEMM that produced `raw`."""
        net, C = self.net, self.C
        sums, ab, mi = (t[g0:g0 + Gc] for t in st)
        if not have_stats:
            ops.bn_stats(raw, sums, Gc, rows, C)
        rm = net.buf(self.name + ".running_mean") if track else None
        rv = net.buf(self.name + ".running_var") if track else None
        nbt = net.buf(self.name + ".num_batches_tracked") if track else None
        ops.bn_finalize_swish_fwd(raw, sums, net.pview(self.name + ".weight"), net.pview(self.name + ".bias"), ab, mi,
                                  rm, rv, nbt, act, Gc, rows, C, 1e-5, 0.1, repeat)

    def backward(self, raw, ab, mi, dAct, G, rows, alloc, key, unscale):
        st = self.alloc_bwd(G, alloc, key)
        return self.run_bwd(raw, ab, mi, dAct, G, rows, st, 0, unscale)

    def alloc_bwd(self, G, alloc, key):
        return alloc(key + ".sums2", (G, self.C, 2), F32, zero="step"), alloc(key + ".coef", (G, self.C, 4), F32)

    def run_bwd(self, raw, ab, mi, dAct, Gc, rows, st, g0, unscale, padded=None):
        """padded = (buffer [rows / 2^w][2^w + 2][C] with zero borders, w): dRaw goes there instead of over dAct"""
        net, C = self.net, self.C
        sums2, coef = st[0][g0:g0 + Gc], st[1][g0:g0 + Gc]
        ops.bn_swish_bwd_reduce(raw, ab[g0:g0 + Gc], mi[g0:g0 + Gc], dAct, sums2, Gc, rows, C)
        if padded is not None:
            ops.bn_bwd_apply_padded(raw, ab[g0:g0 + Gc], mi[g0:g0 + Gc], sums2, dAct, padded[0], padded[1],
                                    net.pview(self.name + ".weight", True), net.pview(self.name + ".bias", True), Gc, rows, C,
                                    unscale)
            return padded[0]
        ops.bn_bwd_apply(raw, ab[g0:g0 + Gc], mi[g0:g0 + Gc], sums2, dAct, net.pview(self.name + ".weight", True),
                         net.pview(self.name + ".bias", True), coef, Gc, rows, C, unscale)
        return dAct  # now dRaw


# ---------------------------------------------------------------------------------------------
# image encoder (vae.py:197-216, 224-242)
# ---------------------------------------------------------------------------------------------
class EncoderExec(_NetBase):
    def __init__(self, arena, prefix, device, cond_dim=0, head_names=("linear_means", "linear_log_var")):
        """cond_dim > 0: CVAE heads Linear(512 + cond_dim, 256) (vae.py:196, 231-237): the 512 feature
        columns go through the tensor cores, the condition columns through mmdyn_linear_f32_acc.
        head_names: the Linear(512 (+cd), 256) layers reading the trunk feature, concatenated along N
        (the two posterior heads; Regressor: its first out_net layer, models.py:57)."""
        super().__init__(arena, prefix, device)
        self.cd = int(cond_dim)
        self.head_names = tuple(head_names)
        self.HN = 256 * len(self.head_names)
        self.c1 = self._pl(plan.conv1_plan("conv1", self.off("conv_net.0.weight")))
        self.c2 = self._pl(plan.conv_s2_plan("conv2", self.off("conv_net.2.weight"), 32, 64, 32))
        self.c3 = self._pl(plan.conv_s2_plan("conv3", self.off("conv_net.5.weight"), 64, 128, 16))
        self.c4 = self._pl(plan.conv_k4s1p0_plan("conv4", self.off("conv_net.8.weight"), 128, 256, 8))
        self.fc = self._pl(plan.linear_plan("fc", [self.off("fc_net.0.weight")], [self.off("fc_net.0.bias")],
                                            6400, [512], k_perm=plan.nhwc_perm(256, 5, 5)))
        self.heads = self._pl(plan.linear_plan(
            "heads", [self.off(n + ".weight") for n in self.head_names], [self.off(n + ".bias") for n in self.head_names],
            512, [256] * len(self.head_names), ld=512 + self.cd))
        self.bn2, self.bn3, self.bn4 = _BN(self, "conv_net.3", 64), _BN(self, "conv_net.6", 128), _BN(self, "conv_net.9", 256)
        self.c1_wg, c1_idx = plan.conv1_wgrad_plan(self.off("conv_net.0.weight"))
        self.c1_idx = torch.from_numpy(c1_idx).to(device)
        self.extra_wgrads = {"conv1": self.c1_idx}

    def forward(self, x, masks, alloc, key, track=True, cond=None):
        """x: (B,3,64,64) fp32 NCHW; cond: (B, cond_dim) fp32 or None; masks: list of (B,512) fp32 dropout masks or None entries (one per
        pass sharing this trunk evaluation; BN running statistics are updated once per mask, as the
        reference's repeated forward passes would).  Returns a record whose 'heads' entry is
        (len(masks)*B, 512) fp32 = [mu | logvar] per mask."""
        self.refresh()
        B, nm = x.shape[0], len(masks)
        r = {"x": x, "B": B, "masks": masks}
        raw1 = alloc(key + ".raw1", (B, 32, 32, 32), F16)
        act1 = alloc(key + ".act1", (B, 32, 32, 32), F16)
        ops.conv1_fwd(x, self.c1.Wf, raw1, B, act=act1)  # raw output + its Swish in one pass
        def conv_bn(pl, bn, a_in, shape, rows, name):
            # conv -> BatchNorm + Swish; the batch statistics come out of the GEMM epilogue when the batch is a whole
            # number of tiles (else mmdyn_bn_stats reads the raw output once more)
            raw, act = alloc(key + ".raw" + name, shape, F16), alloc(key + ".act" + name, shape, F16)
            st = bn.alloc_fwd(1, alloc, key + ".bn" + name)
            fused = bool(FUSE_STATS and B % plan.tile_images(pl.lp.fwd) == 0)
            _ig(pl, "fwd", a_in, raw, B, stats=(st[0], B) if fused else None)
            bn.run_fwd(raw, act, 1, rows, st, 0, track, nm, have_stats=fused)
            return raw, act, (st[1], st[2])

        raw2, act2, r["bn2"] = conv_bn(self.c2, self.bn2, act1, (B, 16, 16, 64), B * 256, "2")
        raw3, act3, r["bn3"] = conv_bn(self.c3, self.bn3, act2, (B, 8, 8, 128), B * 64, "3")
        raw4, act4, r["bn4"] = conv_bn(self.c4, self.bn4, act3, (B, 5, 5, 256), B * 25, "4")
        fc_raw = alloc(key + ".fc_raw", (B, 512), F32)
        _ig(self.fc, "fwd", act4, fc_raw, B, self.fc.bias, True)
        h = alloc(key + ".h", (nm, B, 512), F16)
        ops.swish_dropout_fwd(fc_raw, masks, h, B, 512)
        heads = alloc(key + ".heads", (nm * B, self.HN), F32)
        _ig(self.heads, "fwd", h, heads, nm * B, self.heads.bias, True)
        if self.cd:
            if cond is None or tuple(cond.shape) != (B, self.cd):
                raise ValueError(f"conditional encoder needs a condition of shape ({B}, {self.cd})")
            for j in range(nm):
                for half, nm_ in enumerate(self.head_names):
                    ops.linear_f32_acc(cond, self.pview(nm_ + ".weight")[512:], heads[j * B:(j + 1) * B, 256 * half:],
                                       B, 256, self.cd, self.cd, 512 + self.cd, self.HN)
            r["cond"] = cond
        r.update(raw1=raw1, act1=act1, raw2=raw2, act2=act2, raw3=raw3, act3=act3, raw4=raw4, act4=act4,
                 fc_raw=fc_raw, h=h, heads=heads)
        return r

    def backward(self, r, d_heads, alloc, key, unscale, in_scale=1.0, gp=None):
        """d_heads: (n_masks*B, 512) fp32 gradient; in_scale * d_heads is what flows through the
        fp16 backward (times grad_scale), unscale * in_scale brings parameter gradients back.
        Accumulates parameter gradients into the arena."""
        arena, B, nm = self.arena, r["B"], len(r["masks"])
        rows = nm * B
        HN = self.HN
        dh16 = alloc(key + ".dheads16", (rows, HN), F16)
        ops.f32_to_f16(d_heads, dh16, rows * HN, in_scale)
        db = alloc(key + ".db512", (HN,), F32, zero="step")
        ops.colsum_f32(d_heads, db, rows, HN, HN, unscale * in_scale)
        ops.unpack_add_f32(db, self.heads.bias_idx, arena.grad)
        _wgrad_into(self.heads, r["h"], dh16, rows, arena, alloc, key + ".dW_heads", unscale, gp)
        if self.cd:
            for j in range(nm):
                for half, nm_ in enumerate(self.head_names):
                    ops.linear_f32_wgrad(r["cond"], d_heads[j * B:(j + 1) * B, 256 * half:],
                                         self.pview(nm_ + ".weight", True)[512:], B, 256, self.cd, self.cd,
                                         HN, 512 + self.cd, unscale * in_scale)
        dH = alloc(key + ".dH", (rows, 512), F32)
        _ig(self.heads, "dgrad", dh16, dH, rows, None, True)
        dfc = alloc(key + ".dfc16", (B, 512), F16)
        ops.swish_dropout_bwd(r["fc_raw"], r["masks"], dH, dfc, B, 512)
        ops.colsum_f16(dfc, self.pview("fc_net.0.bias", True), B, 512, 512, unscale)
        _wgrad_into(self.fc, r["act4"], dfc, B, arena, alloc, key + ".dW_fc", unscale, gp)
        d4 = alloc(key + ".d4", (B, 5, 5, 256), F16)
        _ig(self.fc, "dgrad", dfc, d4, B)
        self.bn4.backward(r["raw4"], r["bn4"][0], r["bn4"][1], d4, 1, B * 25, alloc, key + ".bn4", unscale)
        _wgrad_into(self.c4, r["act3"], d4, B, arena, alloc, key + ".dW_c4", unscale, gp)
        d3 = alloc(key + ".d3", (B, 8, 8, 128), F16)
        _ig(self.c4, "dgrad", d4, d3, B)
        self.bn3.backward(r["raw3"], r["bn3"][0], r["bn3"][1], d3, 1, B * 64, alloc, key + ".bn3", unscale)
        _wgrad_into(self.c3, r["act2"], d3, B, arena, alloc, key + ".dW_c3", unscale, gp)
        d2 = alloc(key + ".d2", (B, 16, 16, 64), F16)
        _ig(self.c3, "dgrad", d3, d2, B)
        self.bn2.backward(r["raw2"], r["bn2"][0], r["bn2"][1], d2, 1, B * 256, alloc, key + ".bn2", unscale)
        _wgrad_into(self.c2, r["act1"], d2, B, arena, alloc, key + ".dW_c2", unscale, gp)
        d1 = alloc(key + ".d1", (B, 32, 32, 32), F16)
        _ig(self.c2, "dgrad", d2, d1, B)
        ops.bn_swish_bwd_reduce(r["raw1"], None, None, d1, None, 1, B * 1024, 32)
        # conv1 weight gradient on the tensor cores: x -> NHWC fp16 with 4 channels per pixel and a zero border
        # (written once: the workspace hands out zeroed buffers and the packer never touches the border)
        x8 = alloc(key + ".x4p", (B, 66, 66, plan.LOGIT_CP), F16, zero=not isinstance(alloc, Workspace))
        ops.logit_grad_pack(r["x"], x8, 1.0, B, 64, 64, 1, cp=plan.LOGIT_CP)
        dW1 = gp.view("conv1") if gp is not None else alloc(key + ".dW_c1", tuple(self.c1_idx.shape), F32, zero=True)
        with _on_side(gp):
            ops.wgrad(self.c1_wg, x8, d1, dW1, B, scale=unscale, row_splits=plan.choose_row_splits(self.c1_wg, B),
                      tag="conv1.wgrad", macs_per_img=1024 * 32 * 48)
        if gp is None:
            ops.unpack_add_f32(dW1, self.c1_idx, arena.grad)


# ---------------------------------------------------------------------------------------------
# image decoder (vae.py:263-279, 293-296)
# ---------------------------------------------------------------------------------------------
class DecoderExec(_NetBase):
    def __init__(self, arena, prefix, device, cond_dim=0):
        """cond_dim > 0: CVAE upsample Linear(256 + cond_dim, 6400) (vae.py:257, 286-291): the latent
        columns go through the tensor cores, the condition term is added by mmdyn_cond_add_f16."""
        super().__init__(arena, prefix, device)
        self.cd = int(cond_dim)
        self.up = self._pl(plan.linear_plan("up", [self.off("upsample.0.weight")], [self.off("upsample.0.bias")],
                                            256, [6400], n_perm=plan.nhwc_perm(256, 5, 5), ld=256 + self.cd))
        self.up_rows = torch.from_numpy(np.ascontiguousarray(plan.nhwc_perm(256, 5, 5)).astype(np.int32)).to(device)
        self.d1 = self._pl(plan.deconv_k4s1p0_plan("deconv1", self.off("hallucinate.0.weight"), 256, 128, 5))
        self.d2 = self._pl(plan.deconv_s2_plan("deconv2", self.off("hallucinate.3.weight"), 128, 64, 8))
        lp3 = plan.deconv_s2_plan("deconv3", self.off("hallucinate.6.weight"), 64, 32, 16)
        self.pair_grads = bool(PAIR_GRADS and "pair_dgrad" in lp3.extra)
        if self.pair_grads:  # the backward reads deconv3's output gradient as 128-byte pixel pairs (same packed weights)
            lp3.dgrad, lp3.wgrad = lp3.extra["pair_dgrad"], lp3.extra["pair_wgrad"]
        self.d3 = self._pl(lp3)
        self.d4 = self._pl(plan.deconv_out_plan("deconv4", self.off("hallucinate.9.weight"), 32, 3, 32))
        self.bn1, self.bn2, self.bn3 = _BN(self, "hallucinate.1", 128), _BN(self, "hallucinate.4", 64), _BN(self, "hallucinate.7", 32)

    def forward(self, zh, G, B, alloc, key, track=True, fused_loss=None, cond=None):
        """zh: (G*B, 256) fp16 latent rows, group-major.  Returns record with fp32 NCHW logits.

        fused_loss: dict(target (B,3,64,64), mask|None, dlogits (G*B,66,66,4)|None, loss (fp32 vector),
        slots [loss index per group or -1], gscale, logit_groups (g_lo, g_hi)): the BCE reconstruction
        loss and its logit gradient are computed in the epilogue of the logits layer (problems.py:409-413,
        431-449); logits are then only written for the groups in logit_groups.

        BatchNorm statistics are per group, so the groups are independent and the layer chain can run
        on a subset of the groups at a time (MMDYN_GROUP_CHUNK, see _group_chunk: an L2-residency
        experiment that measured slower than one launch over all groups, which stays the default)."""
        self.refresh()
        R = G * B
        r = {"zh": zh, "G": G, "B": B}
        raw0 = alloc(key + ".raw0", (R, 5, 5, 256), F16)
        act0 = alloc(key + ".act0", (R, 5, 5, 256), F16)
        raw1 = alloc(key + ".raw1", (R, 8, 8, 128), F16)
        act1 = alloc(key + ".act1", (R, 8, 8, 128), F16)
        raw2 = alloc(key + ".raw2", (R, 16, 16, 64), F16)
        act2 = alloc(key + ".act2", (R, 16, 16, 64), F16)
        raw3 = alloc(key + ".raw3", (R, 32, 32, 32), F16)
        act3 = alloc(key + ".act3", (R, 32, 32, 32), F16)
        logits = alloc(key + ".logits", (R, 3, 64, 64), F32)
        s1, s2, s3 = (bn.alloc_fwd(G, alloc, key + nm) for bn, nm in ((self.bn1, ".bn1"), (self.bn2, ".bn2"), (self.bn3, ".bn3")))
        Gc = _group_chunk(G, B)
        cond_rep = None
        if self.cd:
            if cond is None or tuple(cond.shape) != (B, self.cd):
                raise ValueError(f"conditional decoder needs a condition of shape ({B}, {self.cd})")
            cond_rep = alloc(key + ".cond", (R, self.cd), F32)
            cond_rep.view(G, B, self.cd).copy_(cond.unsqueeze(0).expand(G, B, self.cd))  # every group sees the same c
            r["cond_rep"] = cond_rep
        for g0 in range(0, G, Gc):
            sl, n = slice(g0 * B, (g0 + Gc) * B), Gc * B
            _ig(self.up, "fwd", zh[sl], raw0[sl], n, self.up.bias)
            if self.cd:
                ops.cond_add_f16(raw0[sl], cond_rep[sl], self.pview("upsample.0.weight"), self.up_rows, n, 6400,
                                 256 + self.cd, 256, self.cd)
            ops.bn_swish_fwd(raw0[sl], None, act0[sl], 1, n * 25, 256)
            f1 = bool(self.fuse_stats and B % plan.tile_images(self.d1.lp.fwd) == 0)
            _ig(self.d1, "fwd", act0[sl], raw1[sl], n, stats=(s1[0][g0:g0 + Gc], B) if f1 else None)
            self.bn1.run_fwd(raw1[sl], act1[sl], Gc, B * 64, s1, g0, track, have_stats=f1)
            # BatchNorm statistics of raw2 / raw3 come out of the producing GEMM's epilogue (patch kernel);
            # a tile of the 8x8 layer holds two images, which must belong to one group
            f2 = bool(self.fuse_stats and self.d2.lp.fwd.patch and B % 2 == 0)
            f3 = bool(self.fuse_stats and self.d3.lp.fwd.patch)
            _ig(self.d2, "fwd", act1[sl], raw2[sl], n, stats=(s2[0][g0:g0 + Gc], B) if f2 else None)
            self.bn2.run_fwd(raw2[sl], act2[sl], Gc, B * 256, s2, g0, track, have_stats=f2)
            _ig(self.d3, "fwd", act2[sl], raw3[sl], n, stats=(s3[0][g0:g0 + Gc], B) if f3 else None)
            self.bn3.run_fwd(raw3[sl], act3[sl], Gc, B * 1024, s3, g0, track, have_stats=f3)
            if fused_loss is not None:
                fl = fused_loss
                lo, hi = fl["logit_groups"]
                bce = dict(target=fl["target"], mask=fl.get("mask"), loss=fl["loss"], gscale=fl["gscale"],
                           dlogits=fl["dlogits"][sl] if fl.get("dlogits") is not None else None, rows_per_group=B,
                           slots=fl["slots"][g0:g0 + Gc],
                           logit_rows=(max(lo - g0, 0) * B, max(min(hi - g0, Gc), 0) * B))
                _ig(self.d4, "fwd", act3[sl], logits[sl], n, bce=bce)
            else:
                _ig(self.d4, "fwd", act3[sl], logits[sl], n)
            if self.after_group is not None:
                self.after_group(g0, Gc, logits[sl])
        r["bn1"], r["bn2"], r["bn3"] = (s1[1], s1[2]), (s2[1], s2[2]), (s3[1], s3[2])
        r.update(raw0=raw0, act0=act0, raw1=raw1, act1=act1, raw2=raw2, act2=act2, raw3=raw3, act3=act3,
                 logits=logits)
        return r

    after_group = None  # optional callback(g0, Gc, logits_rows) run right after a group chunk's logits (fused losses)
    fuse_stats = FUSE_STATS  # BatchNorm sums in the deconv1/2/3 epilogues

    def backward(self, r, dl8, alloc, key, unscale, gp=None):
        """dl8: (G*B, 66, 66, plan.LOGIT_CP = 4) fp16 logit gradients with a one-pixel ZERO border (3 channels used, times grad_scale).
        Returns dz (G*B, 256) fp32 (times grad_scale)."""
        arena, G, B = self.arena, r["G"], r["B"]
        R = G * B
        g3 = alloc(key + ".g3", (R, 32, 32, 32), F16)
        # deconv3's output gradient as 128-byte pixel pairs: image rows with one zero pixel on each side (zeroed once: the
        # workspace hands out zeroed buffers, the BatchNorm backward writes the 32 inner pixels only)
        g3p = alloc(key + ".g3p", (R, 32, 34, 32), F16, zero=not isinstance(alloc, Workspace)) if self.pair_grads else None
        g2 = alloc(key + ".g2", (R, 16, 16, 64), F16)
        g1 = alloc(key + ".g1", (R, 8, 8, 128), F16)
        g0_ = alloc(key + ".g0", (R, 5, 5, 256), F16)
        b1, b2, b3 = (bn.alloc_bwd(G, alloc, key + nm) for bn, nm in ((self.bn1, ".bn1"), (self.bn2, ".bn2"), (self.bn3, ".bn3")))
        Gc = _group_chunk(G, B)
        for g0 in range(0, G, Gc):
            sl, n = slice(g0 * B, (g0 + Gc) * B), Gc * B
            _wgrad_into(self.d4, dl8[sl], r["act3"][sl], n, arena, alloc, key + ".dW_d4", unscale, gp)
            _ig(self.d4, "dgrad", dl8[sl], g3[sl], n)
            d3g = self.bn3.run_bwd(r["raw3"][sl], r["bn3"][0], r["bn3"][1], g3[sl], Gc, B * 1024, b3, g0, unscale,
                                   padded=(g3p[sl], 5) if self.pair_grads else None)
            _wgrad_into(self.d3, d3g, r["act2"][sl], n, arena, alloc, key + ".dW_d3", unscale, gp)
            _ig(self.d3, "dgrad", d3g, g2[sl], n)
            self.bn2.run_bwd(r["raw2"][sl], r["bn2"][0], r["bn2"][1], g2[sl], Gc, B * 256, b2, g0, unscale)
            _wgrad_into(self.d2, g2[sl], r["act1"][sl], n, arena, alloc, key + ".dW_d2", unscale, gp)
            _ig(self.d2, "dgrad", g2[sl], g1[sl], n)
            self.bn1.run_bwd(r["raw1"][sl], r["bn1"][0], r["bn1"][1], g1[sl], Gc, B * 64, b1, g0, unscale)
            _wgrad_into(self.d1, g1[sl], r["act0"][sl], n, arena, alloc, key + ".dW_d1", unscale, gp)
            _ig(self.d1, "dgrad", g1[sl], g0_[sl], n)
            ops.bn_swish_bwd_reduce(r["raw0"][sl], None, None, g0_[sl], None, 1, n * 25, 256)
        dbp = alloc(key + ".db_up", (6400,), F32, zero="step")
        ops.colsum_f16(g0_, dbp, R, 6400, 6400, unscale)
        ops.unpack_add_f32(dbp, self.up.bias_idx, arena.grad)
        _wgrad_into(self.up, r["zh"], g0_, R, arena, alloc, key + ".dW_up", unscale, gp)
        if self.cd:
            ops.cond_wgrad_f16(g0_, r["cond_rep"], self.pview("upsample.0.weight", True), self.up_rows, R, 6400,
                               256 + self.cd, 256, self.cd, unscale)
        dz = alloc(key + ".dz", (R, 256), F32)
        _ig(self.up, "dgrad", g0_, dz, R, None, True)
        return dz


def _group_chunk(G, B):
    """Groups per launch of the decoder chain.  MMDYN_GROUP_CHUNK overrides (0 = all groups in one
    launch); default: all groups in one launch."""
    env = os.environ.get("MMDYN_GROUP_CHUNK")
    if env is not None:
        v = int(env)
        return G if v <= 0 else max(1, min(G, v))
    return G  # measured (round 1, batch 1024): per-group launches are slower — see DESIGN.md, dead ends


# ---------------------------------------------------------------------------------------------
# pose MLP expert (vae.py:118-123, 219-222, 282-283)
# ---------------------------------------------------------------------------------------------
class PoseExec(_NetBase):
    """Pose expert Linear(7,512)-ReLU-Linear(512,512)-{Linear(512,256)}x2 and Linear(256,512)-ReLU-Linear(512,512)-
    ReLU-Linear(512,7).  The four 256/512-wide layers (99 % of its FLOPs) run on the tensor cores at fp32 accuracy
    through the two-term fp16 split (plan.linear_split_plan, mmdyn_split_f16): forward, dgrad and wgrad are the
    tcgen05 kernels of the image layers with the contraction tripled; the 7-wide edge layers stay on the fp32 SIMT
    kernels (mmdyn_linear_f32_*).  MMDYN_POSE_F32=1 keeps every layer on the SIMT kernels (A/B, debugging)."""

    def __init__(self, arena, device):
        super().__init__(arena, "", device)
        self.tc = os.environ.get("MMDYN_POSE_F32") is None
        o = self.arena.offset
        if self.tc:
            self.e2 = self._pl(plan.linear_split_plan("pose.fc2", [o["pose_encoder.fc_net.2.weight"]],
                                                      [o["pose_encoder.fc_net.2.bias"]], 512, [512]))
            self.eh = self._pl(plan.linear_split_plan(
                "pose.heads", [o["pose_encoder.linear_means.weight"], o["pose_encoder.linear_log_var.weight"]],
                [o["pose_encoder.linear_means.bias"], o["pose_encoder.linear_log_var.bias"]], 512, [256, 256]))
            self.d0 = self._pl(plan.linear_split_plan("pose.dec0", [o["pose_decoder.deconv_net.0.weight"]],
                                                      [o["pose_decoder.deconv_net.0.bias"]], 256, [512]))
            self.d2 = self._pl(plan.linear_split_plan("pose.dec2", [o["pose_decoder.deconv_net.2.weight"]],
                                                      [o["pose_decoder.deconv_net.2.bias"]], 512, [512]))

    def p(self, name, grad=False):
        return self.arena.view(name, self.arena.grad if grad else None)

    # -- split-GEMM building blocks ------------------------------------------------------------
    def _fwd(self, pl, xs, out, M):
        """out[M][N] fp32 = x W^T + b from the split rows xs [M][3K]."""
        _ig(pl, "fwd", xs, out, M, pl.bias, True)

    def _bwd(self, pl, xs, ds, dx, M, unscale, names):
        """dW (+=, straight into the gradient arena, torch layout) and dx [M][K] fp32 from the split rows
        xs [M][3K] (mode 0) and ds [M][3N] (mode 1)."""
        col = 0
        for wg, nm, n_i in zip(pl.lp.extra["wgrads"], names, pl.lp.extra["Ns"]):
            ops.wgrad(wg, xs, ds[:, col:] if col else ds, self.p(nm + ".weight", True).view(n_i, pl.lp.extra["K"]), 3 * M,
                      scale=unscale, row_splits=plan.choose_row_splits(wg, 3 * M), tag=f"{pl.lp.name}.wgrad",
                      macs_per_img=pl.lp.extra["K"] * n_i)
            col += n_i
        if dx is not None:
            _ig(pl, "dgrad", ds, dx, M, None, True)

    # -- encoder -----------------------------------------------------------------------------------
    def enc_forward(self, pose, alloc, key):
        if not self.tc:
            return self._enc_forward_f32(pose, alloc, key)
        self.refresh()
        B = pose.shape[0]
        h1 = alloc(key + ".h1", (B, 512), F32)
        ops.linear_f32_fwd(pose, self.p("pose_encoder.fc_net.0.weight"), self.p("pose_encoder.fc_net.0.bias"), h1,
                           B, 512, 7, 7, 512, 1)
        h1s = alloc(key + ".h1s", (B, 1536), F16)
        ops.split_f16(h1, h1s, B, 512, 0)
        h2 = alloc(key + ".h2", (B, 512), F32)
        self._fwd(self.e2, h1s, h2, B)
        h2s = alloc(key + ".h2s", (B, 1536), F16)
        ops.split_f16(h2, h2s, B, 512, 0)
        heads = alloc(key + ".heads", (B, LATENT_HEADS), F32)
        self._fwd(self.eh, h2s, heads, B)
        return {"pose": pose, "h1": h1, "h1s": h1s, "h2": h2, "h2s": h2s, "heads": heads, "B": B}

    def enc_backward(self, r, d_heads, alloc, key, unscale):
        if not self.tc:
            return self._enc_backward_f32(r, d_heads, alloc, key, unscale)
        B = r["B"]
        dhs = alloc(key + ".dhs", (B, 1536), F16)
        ops.split_f16(d_heads, dhs, B, 512, 1, colsum0=self.p("pose_encoder.linear_means.bias", True),
                      colsum1=self.p("pose_encoder.linear_log_var.bias", True), n_split=256, colsum_scale=unscale)
        dh2 = alloc(key + ".dh2", (B, 512), F32)
        self._bwd(self.eh, r["h2s"], dhs, dh2, B, unscale, ("pose_encoder.linear_means", "pose_encoder.linear_log_var"))
        d2s = alloc(key + ".d2s", (B, 1536), F16)  # fc_net.2 carries no activation (vae.py:17)
        ops.split_f16(dh2, d2s, B, 512, 1, colsum0=self.p("pose_encoder.fc_net.2.bias", True), colsum_scale=unscale)
        dh1 = alloc(key + ".dh1", (B, 512), F32)
        self._bwd(self.e2, r["h1s"], d2s, dh1, B, unscale, ("pose_encoder.fc_net.2",))
        scr = alloc(key + ".scr", (B, 512), F32)
        ops.linear_f32_bwd(r["pose"], self.p("pose_encoder.fc_net.0.weight"), r["h1"], dh1, scr, None,
                           self.p("pose_encoder.fc_net.0.weight", True), self.p("pose_encoder.fc_net.0.bias", True),
                           B, 512, 7, 7, 512, 7, 1, False, unscale)

    # -- decoder -----------------------------------------------------------------------------------
    def dec_forward(self, z, alloc, key):
        if not self.tc:
            return self._dec_forward_f32(z, alloc, key)
        self.refresh()
        R = z.shape[0]
        zs = alloc(key + ".zs", (R, 768), F16)
        ops.split_f16(z, zs, R, 256, 0)
        a1 = alloc(key + ".a1", (R, 512), F32)
        self._fwd(self.d0, zs, a1, R)
        a1s = alloc(key + ".a1s", (R, 1536), F16)
        ops.split_f16(a1, a1s, R, 512, 0, relu=True, x_out=a1)   # a1 := relu(a1), in place
        a2 = alloc(key + ".a2", (R, 512), F32)
        self._fwd(self.d2, a1s, a2, R)
        ops.relu_f32(a2, a2)
        rec = alloc(key + ".rec", (R, 7), F32)
        ops.linear_f32_fwd(a2, self.p("pose_decoder.deconv_net.4.weight"), self.p("pose_decoder.deconv_net.4.bias"),
                           rec, R, 7, 512, 512, 7, 0)
        return {"z": z, "zs": zs, "a1": a1, "a1s": a1s, "a2": a2, "rec": rec, "R": R}

    def dec_backward(self, r, d_rec, alloc, key, unscale):
        if not self.tc:
            return self._dec_backward_f32(r, d_rec, alloc, key, unscale)
        R = r["R"]
        scr = alloc(key + ".scr", (R, 512), F32)
        da2 = alloc(key + ".da2", (R, 512), F32)
        ops.linear_f32_bwd(r["a2"], self.p("pose_decoder.deconv_net.4.weight"), r["rec"], d_rec, scr, da2,
                           self.p("pose_decoder.deconv_net.4.weight", True),
                           self.p("pose_decoder.deconv_net.4.bias", True), R, 7, 512, 512, 7, 512, 0, False, unscale)
        d2s = alloc(key + ".d2s", (R, 1536), F16)
        ops.split_f16(da2, d2s, R, 512, 1, mask_y=r["a2"], colsum0=self.p("pose_decoder.deconv_net.2.bias", True),
                      colsum_scale=unscale)
        da1 = alloc(key + ".da1", (R, 512), F32)
        self._bwd(self.d2, r["a1s"], d2s, da1, R, unscale, ("pose_decoder.deconv_net.2",))
        d1s = alloc(key + ".d1s", (R, 1536), F16)
        ops.split_f16(da1, d1s, R, 512, 1, mask_y=r["a1"], colsum0=self.p("pose_decoder.deconv_net.0.bias", True),
                      colsum_scale=unscale)
        dz = alloc(key + ".dz", (R, 256), F32)
        self._bwd(self.d0, r["zs"], d1s, dz, R, unscale, ("pose_decoder.deconv_net.0",))
        return dz

    # -- all-SIMT fp32 path (MMDYN_POSE_F32=1) ---------------------------------------------------
    def _enc_forward_f32(self, pose, alloc, key):
        B = pose.shape[0]
        h1 = alloc(key + ".h1", (B, 512), F32)
        ops.linear_f32_fwd(pose, self.p("pose_encoder.fc_net.0.weight"), self.p("pose_encoder.fc_net.0.bias"), h1,
                           B, 512, 7, 7, 512, 1)
        h2 = alloc(key + ".h2", (B, 512), F32)
        ops.linear_f32_fwd(h1, self.p("pose_encoder.fc_net.2.weight"), self.p("pose_encoder.fc_net.2.bias"), h2,
                           B, 512, 512, 512, 512, 0)
        heads = alloc(key + ".heads", (B, LATENT_HEADS), F32)
        ops.linear_f32_fwd(h2, self.p("pose_encoder.linear_means.weight"), self.p("pose_encoder.linear_means.bias"),
                           heads, B, 256, 512, 512, 512, 0)
        ops.linear_f32_fwd(h2, self.p("pose_encoder.linear_log_var.weight"),
                           self.p("pose_encoder.linear_log_var.bias"), heads[:, 256:], B, 256, 512, 512, 512, 0)
        return {"pose": pose, "h1": h1, "h2": h2, "heads": heads, "B": B}

    def _enc_backward_f32(self, r, d_heads, alloc, key, unscale):
        B = r["B"]
        scr = alloc(key + ".scr", (B, 512), F32)
        dh2 = alloc(key + ".dh2", (B, 512), F32)
        for i, nm in enumerate(("linear_means", "linear_log_var")):
            ops.linear_f32_bwd(r["h2"], self.p(f"pose_encoder.{nm}.weight"), r["heads"][:, 256 * i:],
                               d_heads[:, 256 * i:], scr, dh2, self.p(f"pose_encoder.{nm}.weight", True),
                               self.p(f"pose_encoder.{nm}.bias", True), B, 256, 512, 512, 512, 512, 0, i == 1, unscale)
        dh1 = alloc(key + ".dh1", (B, 512), F32)
        ops.linear_f32_bwd(r["h1"], self.p("pose_encoder.fc_net.2.weight"), r["h2"], dh2, scr, dh1,
                           self.p("pose_encoder.fc_net.2.weight", True), self.p("pose_encoder.fc_net.2.bias", True),
                           B, 512, 512, 512, 512, 512, 0, False, unscale)
        ops.linear_f32_bwd(r["pose"], self.p("pose_encoder.fc_net.0.weight"), r["h1"], dh1, scr, None,
                           self.p("pose_encoder.fc_net.0.weight", True), self.p("pose_encoder.fc_net.0.bias", True),
                           B, 512, 7, 7, 512, 7, 1, False, unscale)

    def _dec_forward_f32(self, z, alloc, key):
        R = z.shape[0]
        a1 = alloc(key + ".a1", (R, 512), F32)
        ops.linear_f32_fwd(z, self.p("pose_decoder.deconv_net.0.weight"), self.p("pose_decoder.deconv_net.0.bias"),
                           a1, R, 512, 256, 256, 512, 1)
        a2 = alloc(key + ".a2", (R, 512), F32)
        ops.linear_f32_fwd(a1, self.p("pose_decoder.deconv_net.2.weight"), self.p("pose_decoder.deconv_net.2.bias"),
                           a2, R, 512, 512, 512, 512, 1)
        rec = alloc(key + ".rec", (R, 7), F32)
        ops.linear_f32_fwd(a2, self.p("pose_decoder.deconv_net.4.weight"), self.p("pose_decoder.deconv_net.4.bias"),
                           rec, R, 7, 512, 512, 7, 0)
        return {"z": z, "a1": a1, "a2": a2, "rec": rec, "R": R}

    def _dec_backward_f32(self, r, d_rec, alloc, key, unscale):
        R = r["R"]
        scr = alloc(key + ".scr", (R, 512), F32)
        da2 = alloc(key + ".da2", (R, 512), F32)
        ops.linear_f32_bwd(r["a2"], self.p("pose_decoder.deconv_net.4.weight"), r["rec"], d_rec, scr, da2,
                           self.p("pose_decoder.deconv_net.4.weight", True),
                           self.p("pose_decoder.deconv_net.4.bias", True), R, 7, 512, 512, 7, 512, 0, False, unscale)
        da1 = alloc(key + ".da1", (R, 512), F32)
        ops.linear_f32_bwd(r["a1"], self.p("pose_decoder.deconv_net.2.weight"), r["a2"], da2, scr, da1,
                           self.p("pose_decoder.deconv_net.2.weight", True),
                           self.p("pose_decoder.deconv_net.2.bias", True), R, 512, 512, 512, 512, 512, 1, False, unscale)
        dz = alloc(key + ".dz", (R, 256), F32)
        ops.linear_f32_bwd(r["z"], self.p("pose_decoder.deconv_net.0.weight"), r["a1"], da1, scr, dz,
                           self.p("pose_decoder.deconv_net.0.weight", True),
                           self.p("pose_decoder.deconv_net.0.bias", True), R, 512, 256, 256, 512, 256, 1, False, unscale)
        return dz


# ---------------------------------------------------------------------------------------------
# pose regressor baseline (models.py:28-77)
# ---------------------------------------------------------------------------------------------
class RegressorExec:
    """Regressor.forward: the image-encoder trunk (same conv_net / fc_net as the cnn Encoder) followed by
    out_net = Linear(512 (+cd), 256) -> ReLU -> Linear(256, 256) -> ReLU -> Linear(256, out_dim).  The trunk
    and out_net.0 run on the EncoderExec kernels (out_net.0 takes the place of the posterior heads,
    condition columns included), the two small Linears behind it in fp32 on the pose-MLP kernels."""

    def __init__(self, arena, device, cond_dim=0):
        self.arena, self.device = arena, device
        self.trunk = EncoderExec(arena, "", device, cond_dim, head_names=("out_net.0",))
        self.out_dim = int(arena.view("out_net.4.bias").numel())

    def p(self, name, grad=False):
        return self.arena.view(name, self.arena.grad if grad else None)

    def forward(self, x, mask, alloc, key, track=True, cond=None):
        B = x.shape[0]
        r = self.trunk.forward(x, [mask], alloc, key, track, cond)
        a1 = alloc(key + ".a1", (B, 256), F32)
        ops.relu_f32(r["heads"], a1)
        a2 = alloc(key + ".a2", (B, 256), F32)
        ops.linear_f32_fwd(a1, self.p("out_net.2.weight"), self.p("out_net.2.bias"), a2, B, 256, 256, 256, 256, 1)
        out = alloc(key + ".out", (B, self.out_dim), F32)
        ops.linear_f32_fwd(a2, self.p("out_net.4.weight"), self.p("out_net.4.bias"), out, B, self.out_dim, 256, 256,
                           self.out_dim, 0)
        r.update(a1=a1, a2=a2, out=out)
        return r

    def backward(self, r, d_out, alloc, key, unscale=1.0):
        """d_out: (B, out_dim) fp32.  Parameter gradients accumulate into the arena."""
        B, od = r["B"], self.out_dim
        scr = alloc(key + ".scr", (B, 256), F32)
        da2 = alloc(key + ".da2", (B, 256), F32)
        ops.linear_f32_bwd(r["a2"], self.p("out_net.4.weight"), r["out"], d_out, scr, da2,
                           self.p("out_net.4.weight", True), self.p("out_net.4.bias", True), B, od, 256, 256, od, 256,
                           0, False, unscale)
        da1 = alloc(key + ".da1", (B, 256), F32)
        ops.linear_f32_bwd(r["a1"], self.p("out_net.2.weight"), r["a2"], da2, scr, da1,
                           self.p("out_net.2.weight", True), self.p("out_net.2.bias", True), B, 256, 256, 256, 256, 256,
                           1, False, unscale)
        d_heads = alloc(key + ".d_heads", (B, 256), F32)
        ops.act_grad_f32(r["a1"], da1, d_heads, B, 256, 256, 1)
        # the fp16 trunk backward carries gradients multiplied by gs = B (like the ELBO step, whose loss is
        # divided by B): a 'sum'-reduced MSE gradient is O(1) per sample already, so in_scale = 1
        self.trunk.backward(r, d_heads, alloc, key, unscale, 1.0)


def get_execs(module, device):
    """Build (once per module and device) the executors of every sub-network present."""
    arena = get_arena(module, device)
    ex = module.__dict__.get("_mmdyn_execs")
    if ex is None or ex["device"] != arena.flat.device or ex["flat_id"] != id(arena.flat):
        dev = arena.flat.device
        ex = {"device": dev, "flat_id": id(arena.flat), "enc": {}, "dec": {}, "pose": None}
        names = set(n.split(".")[0] for n in arena.names)
        cd = int(getattr(module, "condition_dim", 0) or 0) if getattr(module, "conditional", False) else 0
        ex["cond_dim"] = cd
        for n in sorted(names):
            if n.endswith("encoder") and n != "pose_encoder":
                ex["enc"][n] = EncoderExec(arena, n, dev, cd)
            elif n.endswith("decoder") and n != "pose_decoder":
                ex["dec"][n] = DecoderExec(arena, n, dev, cd)
        if "pose_encoder" in names:
            ex["pose"] = PoseExec(arena, dev)
        if "out_net" in names and "conv_net" in names:  # Regressor (models.py:28-77): the module IS the trunk
            ex["reg"] = RegressorExec(arena, dev, int(getattr(module, "num_classes", 0) or 0)
                                      if getattr(module, "conditional", False) else 0)
            ex["enc"]["__regressor__"] = ex["reg"].trunk
        nets = list(ex["enc"].values()) + list(ex["dec"].values())
        packed = nets + ([ex["pose"]] if ex["pose"] is not None else [])  # the pose expert's split weights are packed too
        ex["packer"] = ModelPacker(packed, dev)
        for net in packed:
            net.packer = ex["packer"]
        ex["gradpack"] = {net.prefix: GradPack(net, dev) for net in nets}
        module.__dict__["_mmdyn_execs"] = ex
    return arena, ex


# ---------------------------------------------------------------------------------------------
# fused training / evaluation step
# ---------------------------------------------------------------------------------------------
class _StepLossFn(torch.autograd.Function):
    """Connects the fused step to torch autograd: loss.backward() (problems.py:153) runs the
    hand-written backward of the whole step into the gradient arena."""

    @staticmethod
    def forward(ctx, anchor, loss_value, eng, token):
        ctx.eng, ctx.token = eng, token
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.eng._backward(ctx.token, grad_out)
        return None, None, None, None


class StepEngine:
    """One object per model: evaluate() is Reconstruction._evaluate_mvae / the VAE branch of
    SeqModeling._evaluate_model (problems.py:473-546, 702-716) as one fused launch sequence.

    kind 'vae'  : one pass {x}; posterior = encoder output (no prior expert).
    kind 'mvae' : passes {v,t} {v} {t} and, with use_pose, {v,t,p} {v,p} {t,p} {p}; the image
                  encoder trunks are evaluated once and shared; decoders run group-batched over
                  the passes that carry their loss.
    exact_running_stats: also run the decoder invocations whose outputs the reference discards
                  (vae.py:160-161) so BatchNorm running statistics match it bit-for-bit in count.
    """

    IMG = 3 * 64 * 64

    def __init__(self, model, kind, use_pose=False, pose_multiplier=1000.0, noise_src=None,
                 exact_running_stats=False, grad_scale=None):
        self.model, self.kind, self.use_pose = model, kind, bool(use_pose)
        self.pose_multiplier = float(pose_multiplier)
        self.noise_src = noise_src
        self.exact = bool(exact_running_stats)
        self.grad_scale = grad_scale
        # reconstruction BCE + logit gradient in the epilogue of the logits layer (igemm out_mode 5);
        # MMDYN_NO_FUSED_BCE=1 keeps the separate mmdyn_bce_logits pass (A/B measurements, debugging)
        self.fuse_bce = os.environ.get("MMDYN_NO_FUSED_BCE") is None
        if kind == "vae":
            self.mods = {"x": ("encoder", "decoder")}
            self.passes = [("x",)]
            self.use_prior = False
        else:
            self.mods = {"v": ("visual_encoder", "visual_decoder"), "t": ("tactile_encoder", "tactile_decoder")}
            self.passes = [("v", "t"), ("v",), ("t",)]
            if self.use_pose:
                self.passes += [("v", "t", "p"), ("v", "p"), ("t", "p"), ("p",)]
            self.use_prior = True
        self.ws = None
        self._token = 0
        self._state = None
        self.always_refresh = False   # set while capturing a CUDA graph: the packing kernels must be in it
        self.bucket_hook = None       # callable(prefixes) fired when those sub-networks' gradients are final
        # independent sub-networks (visual / tactile branch, pose MLP) run on separate streams: at
        # small batch a single branch cannot fill 148 SMs; in a captured graph these become
        # parallel branches
        self.concurrent = os.environ.get("MMDYN_SERIAL_BRANCHES") is None
        self.stagger = int(os.environ.get("MMDYN_STAGGER", "0"))  # offset between the image branches, in GEMM launches
        self._side = {}

    # -- helpers ------------------------------------------------------------------------------
    def _noise(self):
        if self.noise_src is not None:
            return self.noise_src
        src = getattr(self.model, "noise", None)
        if src is not None:
            return src
        from . import noise
        return noise.get_default()

    def set_concurrent(self, flag):
        """Serialise (False) or parallelise (True) the independent branches — the per-kernel CUDA-event
        profile of bench.py needs them serial, otherwise kernels time-share the SMs."""
        self.concurrent = bool(flag)
        _, ex = get_execs(self.model, None)
        for gp in ex["gradpack"].values():
            if flag and gp.wstream is None:
                gp.wstream = torch.cuda.Stream()
            elif not flag:
                gp.wstream = None

    def _fork(self, branches, main_fn=None):
        """Run `branches` (dict key -> callable) concurrently on per-key side streams, `main_fn` on the
        current stream, then join everything back into the current stream."""
        if not self.concurrent or len(branches) == 0:
            for fn in branches.values():
                fn()
            if main_fn is not None:
                main_fn()
            return
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        used = []
        prev_evt = None
        for k, fn in branches.items():
            st = self._side.get(k)
            if st is None:
                st = self._side[k] = torch.cuda.Stream()
            st.wait_event(ev)
            if prev_evt is not None:
                st.wait_event(prev_evt)  # start this branch behind the previous branch's first GEMM(s): see ops.arm_stagger
            with torch.cuda.stream(st):
                evt = torch.cuda.Event() if self.stagger > 0 else None
                ops.arm_stagger(evt, self.stagger)
                fn()
                ops.arm_stagger(None, 0)
            prev_evt = evt
            used.append(st)
        if main_fn is not None:
            main_fn()
        for st in used:
            cur.wait_stream(st)

    def _setup(self, device):
        arena, ex = get_execs(self.model, device)
        if self.ws is None or self.ws.device != arena.flat.device:
            self.ws = Workspace(arena.flat.device)
        return arena, ex

    def evaluate(self, x, targets, kl_weight, loss_mask=None, want_outputs=True, need_grad=None, autograd=True,
                 condition=None, track=True):
        """x / targets: tensor (vae) or list [visual, tactile(, pose)] (mvae), fp32, on the GPU.
        Returns (outputs, loss) like the reference; loss.backward() then fills the parameter
        gradients.  Under torch.no_grad() only the forward runs (Problem._test_epoch).
        need_grad / autograd=False: keep the backward state without an autograd node, for callers
        that invoke backward() themselves (CUDA-graph capture of the whole step).
        track=False: BatchNorm running statistics / num_batches_tracked are left untouched (the warm-up
        passes of a graph capture must not move model buffers)."""
        if self.kind == "vae":
            xs, ts = {"x": x}, {"x": targets}
        else:
            xs = {"v": x[0], "t": x[1]}
            ts = {"v": targets[0], "t": targets[1]}
            if self.use_pose:
                xs["p"], ts["p"] = x[2], targets[2]
        first = next(iter(xs.values()))
        if not first.is_cuda:
            raise RuntimeError("mmdyn_b200 has no CPU path: inputs must be CUDA tensors")
        if not self.model.training:
            raise NotImplementedError("eval-mode BatchNorm/Dropout is not part of the reference path "
                                      "(the reference keeps model.train() even for validation: problems.py:174)")
        arena, ex = self._setup(first.device)
        ws, B, D = self.ws, first.shape[0], 256
        ws.begin_step()
        cond = None
        if ex["cond_dim"]:  # CVAE: the condition (shock force) enters every image encoder head and decoder
            if condition is None:
                raise ValueError("this model was built with conditional=True: evaluate() needs `condition`")
            cond = condition.float()
            cond = (cond.unsqueeze(1) if cond.dim() == 1 else cond).contiguous()
        for k in xs:
            xs[k] = xs[k].contiguous().float()
            ts[k] = ts[k].contiguous().float()
        need_grad = torch.is_grad_enabled() if need_grad is None else bool(need_grad)
        if self.always_refresh:
            ex["packer"].token = None
        gs = float(self.grad_scale) if self.grad_scale else float(B)
        npass = len(self.passes)
        src = self._noise()

        # groups: which passes each decoder runs (loss-bearing ones, or all with exact stats)
        img_mods = [m for m in self.mods]
        enc_passes = {m: [i for i, p in enumerate(self.passes) if m in p] for m in img_mods}
        dec_groups = {m: (list(range(npass)) if self.exact else enc_passes[m]) for m in img_mods}
        pose_passes = [i for i, p in enumerate(self.passes) if "p" in p]

        # noise in the reference's consumption order: per pass visual mask, tactile mask, eps
        masks = {m: [None] * len(enc_passes[m]) for m in img_mods}
        eps = ws("eps", (npass, B, D), F32)
        mbuf = {m: ws("mask_" + m, (len(enc_passes[m]), B, 512), F32) for m in img_mods}
        if hasattr(src, "fill_step"):
            # device RNG: no ordering constraint between draws -> one launch per buffer
            src.fill_step([mbuf[m] for m in img_mods], eps)
            for m in img_mods:
                masks[m] = [mbuf[m][j] for j in range(len(enc_passes[m]))]
        else:
            for i, p in enumerate(self.passes):
                for m in img_mods:
                    if m in p:
                        j = enc_passes[m].index(i)
                        masks[m][j] = src.dropout_mask(B, first.device, out=mbuf[m][j])
                src.normal(B, D, first.device, out=eps[i])

        # encoders (once per modality)
        # before the fork: every branch reads the packed weights (decoder / data-gradient copies arrive on a side stream)
        # (measured: −0.5 .. −1 % per step at batch >= 256, +2 % at 128 where the side-stream gathers get in the encoders' way)
        ex["packer"].refresh(arena, staged=self.concurrent and STAGED_PACK and B >= 256)
        enc_rec, pose_box = {}, {}

        def enc_branch(m):
            def fn():
                enc_rec[m] = ex["enc"][self.mods[m][0]].forward(xs[m], masks[m], ws, "enc_" + m, track, cond)
            return fn

        def pose_enc():
            pose_box["rec"] = ex["pose"].enc_forward(xs["p"], ws, "penc")
        self._fork({m: enc_branch(m) for m in img_mods}, pose_enc if self.use_pose else None)
        pose_rec = pose_box.get("rec")

        # PoE + reparam + KL per pass
        scal = ws("scal", (64,), F32, zero="step")
        mu_all, lv_all = ws("mu", (npass, B, D), F32), ws("lv", (npass, B, D), F32)
        zscr = ws("z_scratch", (B, D), F32)
        zdec = {m: ws("z_" + m, (len(dec_groups[m]) * B, D), F16) for m in img_mods}
        zpose = ws("z_p", (max(1, len(pose_passes)) * B, D), F32) if self.use_pose else None

        def experts(i):
            out = []
            for m in img_mods:
                if m in self.passes[i]:
                    j = enc_passes[m].index(i)
                    out.append(enc_rec[m]["heads"][j * B:(j + 1) * B])
            if "p" in self.passes[i]:
                out.append(pose_rec["heads"])
            return out

        fwd_passes = []
        for i, p in enumerate(self.passes):
            hs = experts(i)
            zh = []
            for m in img_mods:
                if i in dec_groups[m]:
                    g = dec_groups[m].index(i)
                    zh.append(zdec[m][g * B:(g + 1) * B])
            zf = zpose[pose_passes.index(i) * B:(pose_passes.index(i) + 1) * B] if "p" in p else zscr
            fwd_passes.append(dict(mu_e=[h[:, :D] for h in hs], lv_e=[h[:, D:] for h in hs], eps=eps[i], mu=mu_all[i],
                                   lv=lv_all[i], z=zf, zh=zh[0] if len(zh) > 0 else None,
                                   zh2=zh[1] if len(zh) > 1 else None, kl_sum=scal[i:i + 1]))
        # every sub-sampled pass in one launch (z_scratch is shared by the passes without a pose decoder: nobody reads it)
        ops.poe_fwd_multi(fwd_passes, self.use_prior, 2 * D, B, D)

        # decoders, group-batched, + losses (and logit gradients when training): one branch per modality
        slot, nslot = {}, 8
        for m in img_mods:
            for i in enc_passes[m]:
                slot[(m, i)] = nslot
                nslot += 1
        for i in pose_passes:
            slot[("p", i)] = nslot
            nslot += 1
        dec_rec, dl8, pdec_box = {}, {}, {}

        def dec_branch(m):
            def fn():
                G = len(dec_groups[m])
                dex = ex["dec"][self.mods[m][1]]
                # one-pixel zero border (zeroed at allocation, never written): deconv4's backward reads
                # the 4 x-taps of a pixel as one 64-byte window (plan.deconv_out_plan)
                dl8[m] = ws("dl8_" + m, (G * B, 66, 66, plan.LOGIT_CP), F16, zero=self.exact) if need_grad else None

                slots = [slot[(m, i)] if i in enc_passes[m] else -1 for i in dec_groups[m]]
                if self.fuse_bce:
                    # logits are only materialised where something reads them: the joint pass for
                    # `outputs`, every group when the (unmasked) metric must be recomputed from them
                    if not want_outputs:
                        lg = (0, 0)
                    elif loss_mask is not None or self.kind == "vae":
                        lg = (0, G)
                    else:
                        jg = dec_groups[m].index(3 if self.use_pose else 0)
                        lg = (jg, jg + 1)
                    fl = dict(target=ts[m], mask=loss_mask, dlogits=dl8[m], loss=scal, slots=slots, gscale=gs / B,
                              logit_groups=lg)
                    dec_rec[m] = dex.forward(zdec[m], G, B, ws, "dec_" + m, track, fused_loss=fl, cond=cond)
                    return

                def losses(g0, Gc, lg_rows):  # right after a group chunk's logits, while they are L2-resident
                    for g in range(g0, g0 + Gc):
                        if slots[g] < 0:
                            continue
                        k = slots[g]
                        ops.bce_logits(lg_rows[(g - g0) * B:(g - g0 + 1) * B], ts[m], loss_mask, scal[k:k + 1],
                                       dl8[m][g * B:(g + 1) * B] if need_grad else None, gs / B, B, 64, 64, 1)
                dex.after_group = losses
                try:
                    dec_rec[m] = dex.forward(zdec[m], G, B, ws, "dec_" + m, track, cond=cond)
                finally:
                    dex.after_group = None
            return fn

        def pose_dec():
            pdec_box["rec"] = ex["pose"].dec_forward(zpose, ws, "pdec")
            pdec_box["d"] = ws("d_prec", (len(pose_passes) * B, 7), F32) if need_grad else None
            for g, i in enumerate(pose_passes):
                k = slot[("p", i)]
                ops.mse(pdec_box["rec"]["rec"][g * B:(g + 1) * B], ts["p"], scal[k:k + 1],
                        pdec_box["d"][g * B:(g + 1) * B] if need_grad else None, self.pose_multiplier, gs / B, B * 7)
        ex["packer"].wait_stage(1)
        self._fork({m: dec_branch(m) for m in img_mods}, pose_dec if self.use_pose else None)
        ex["packer"].wait_stage(2)  # joins the pack stream: the backward (and whoever changes the arena next) is ordered behind it
        pdec_rec, d_prec = pdec_box.get("rec"), pdec_box.get("d")
        klw = float(kl_weight)
        loss_value = (scal[8:nslot].sum() + klw * scal[:npass].sum()) / B

        self._token += 1
        self._state = dict(token=self._token, arena=arena, ex=ex, B=B, gs=gs, klw=klw, enc_rec=enc_rec,
                           pose_rec=pose_rec, dec_rec=dec_rec, pdec_rec=pdec_rec, dl8=dl8, d_prec=d_prec,
                           enc_passes=enc_passes, dec_groups=dec_groups, pose_passes=pose_passes, eps=eps,
                           experts=experts, img_mods=img_mods) if need_grad else None
        loss = _StepLossFn.apply(arena.params[0], loss_value, self, self._token) if (need_grad and autograd) \
            else loss_value

        outputs = None
        if want_outputs:
            outputs = self._outputs(dec_rec, pdec_rec, mu_all, lv_all, scal, slot, ts, loss_mask, B,
                                    dec_groups, pose_passes)
        return outputs, loss

    def _outputs(self, dec_rec, pdec_rec, mu_all, lv_all, scal, slot, ts, loss_mask, B, dec_groups, pose_passes):
        """The reference's `outputs` dict (problems.py:537-544, 714-715), including its name-rebinding
        quirks: recon_x = reconstructions of the joint pass, means/log_var = posterior of the LAST pass.
        perf_measure values are 0-dim device tensors (no host sync inside the step)."""
        n_el = float(B * self.IMG)
        if self.kind == "vae":
            lg = dec_rec["x"]["logits"]
            if loss_mask is None:
                meas = scal[slot[("x", 0)]] / n_el
            else:
                tmp = self.ws("metric_tmp", (4,), F32, zero=True)
                ops.bce_logits(lg, ts["x"], None, tmp[0:1], None, 0.0, B, 64, 64)
                meas = tmp[0] / n_el
            return {"recon_x": lg.clone(), "means": mu_all[0].clone(), "log_var": lv_all[0].clone(),
                    "perf_measure": {"x": meas}}
        joint = 3 if self.use_pose else 0
        rec = []
        for m in ("v", "t"):
            g = dec_groups[m].index(joint)
            rec.append(dec_rec[m]["logits"][g * B:(g + 1) * B].clone())
        perf = {}
        for m, name, uni in (("v", "visual", 1), ("t", "tactile", 2)):
            if loss_mask is None:
                perf[name] = scal[slot[(m, uni)]] / n_el
            else:
                tmp = self.ws("metric_tmp_" + m, (4,), F32, zero=True)
                g = dec_groups[m].index(uni)
                ops.bce_logits(dec_rec[m]["logits"][g * B:(g + 1) * B], ts[m], None, tmp[0:1], None, 0.0, B, 64, 64)
                perf[name] = tmp[0] / n_el
        if self.use_pose:
            g = pose_passes.index(joint)
            rec.append(pdec_rec["rec"][g * B:(g + 1) * B].clone())
            perf["pose"] = scal[slot[("p", 6)]] / (self.pose_multiplier * B * 7)
        last = len(self.passes) - 1
        return {"recon_x": rec, "means": mu_all[last].clone(), "log_var": lv_all[last].clone(), "perf_measure": perf}

    def backward(self):
        """Run the fused backward of the last evaluate(need_grad=True) (no autograd involved)."""
        if self._state is None:
            raise RuntimeError("mmdyn_b200: no pending step to differentiate")
        self._backward(self._state["token"])

    def _backward(self, token, grad_out=None):
        st = self._state
        if st is None or st["token"] != token:
            raise RuntimeError("mmdyn_b200: backward() called for a step whose buffers were overwritten by a "
                               "later evaluate(); call loss.backward() before evaluating the next batch")
        # autograd route only (loss.backward()): the hand-written backward is the gradient of `loss` itself,
        # so an upstream gradient other than 1 (e.g. (2 * loss).backward()) must fail loudly instead of
        # producing unscaled gradients.  One 4-byte read-back per eager step; the graph path never comes here.
        if grad_out is not None and os.environ.get("MMDYN_NO_CHECK_GRAD_OUT") is None:
            if abs(float(grad_out) - 1.0) > 1e-6:
                raise RuntimeError("mmdyn_b200: the fused step's backward is d(loss)/d(params) for an upstream "
                                   f"gradient of 1, got {float(grad_out)}; scale the learning rate or the "
                                   "gradients (optimizer.grad_prescale) instead of the loss")
        arena, ex, ws, B, gs = st["arena"], st["ex"], self.ws, st["B"], st["gs"]
        arena.attach_grads()
        unscale = 1.0 / gs
        D = 256
        img_mods = st["img_mods"]
        hook = self.bucket_hook or (lambda prefixes: None)
        dz, dzp_box = {}, {}
        gps = ex["gradpack"]

        def dec_bwd(m):
            def fn():
                gp = gps[self.mods[m][1]]
                gp.begin()
                dz[m] = ex["dec"][self.mods[m][1]].backward(st["dec_rec"][m], st["dl8"][m], ws, "dec_" + m, unscale, gp)
                gp.flush(arena)
                hook([self.mods[m][1]])
            return fn

        def pose_dec_bwd():
            dzp_box["dz"] = ex["pose"].dec_backward(st["pdec_rec"], st["d_prec"], ws, "pdec", unscale)
        self._fork({m: dec_bwd(m) for m in img_mods}, pose_dec_bwd if self.use_pose else None)
        dzp = dzp_box.get("dz")
        dh = {m: ws("dheads_" + m, (len(st["enc_passes"][m]) * B, 512), F32, zero=True) for m in img_mods}
        dhp = ws("dheads_p", (B, 512), F32, zero=True) if self.use_pose else None
        bwd_passes = []
        for i, p in enumerate(self.passes):
            hs = st["experts"](i)
            outs, dzs = [], []
            for m in img_mods:
                if m in p:
                    j = st["enc_passes"][m].index(i)
                    outs.append(dh[m][j * B:(j + 1) * B])
                    g = st["dec_groups"][m].index(i)
                    dzs.append(dz[m][g * B:(g + 1) * B])
            if "p" in p:
                outs.append(dhp)
                g = st["pose_passes"].index(i)
                dzs.append(dzp[g * B:(g + 1) * B])
            bwd_passes.append(dict(mu_e=[h[:, :D] for h in hs], lv_e=[h[:, D:] for h in hs], eps=st["eps"][i], dz=dzs,
                                   dmu_e=[o[:, :D] for o in outs], dlv_e=[o[:, D:] for o in outs]))
        # one launch for all passes; the pose expert's gradient rows are shared by its 4 passes (atomic accumulation)
        ops.poe_bwd_multi(bwd_passes, self.use_prior, 2 * D, st["klw"] * gs / B, 2 * D, True, B, D)
        def enc_bwd(m):
            def fn():
                gp = gps[self.mods[m][0]]
                gp.begin()
                ex["enc"][self.mods[m][0]].backward(st["enc_rec"][m], dh[m], ws, "enc_" + m, unscale, 1.0, gp)
                gp.flush(arena)
                hook([self.mods[m][0]])
            return fn

        def pose_enc_bwd():
            ex["pose"].enc_backward(st["pose_rec"], dhp, ws, "penc", unscale)
            hook(["pose_encoder", "pose_decoder"])
        self._fork({m: enc_bwd(m) for m in img_mods}, pose_enc_bwd if self.use_pose else None)
        self._state = None


class GraphedTrainStep:
    """zero_grad + fused forward + fused backward (+ optimizer) captured once in a CUDA graph and
    replayed per batch: the ~300 kernel launches of a cnn-mvae+pose step cost one graph launch.

    Inputs are copied into static device buffers (`load`), results are static tensors overwritten
    by every replay.  The noise source must be a DeviceNoise (device-resident Philox counter) and
    the optimizer a FusedAdam / FusedSGD (device-resident step counter), so that replays advance.
    With `grad_sync` (data parallel) the graph ends after the backward; the caller all-reduces the
    gradient arena and then calls `apply()`, a second graph holding the optimizer update."""

    def __init__(self, step_engine, optimizer, example_x, example_t, kl_weight, loss_mask=None, split_optimizer=False,
                 warmup=2, grad_sync=None, condition=None, peer_exchange=None):
        """grad_sync: a parallel.GradSync whose bucketed NCCL all-reduces are captured INSIDE the graph
        (launched from the backward on a side stream, overlapping the encoder backward); the
        alternative for data parallelism is split_optimizer=True (backward graph, eager all-reduce,
        optimizer graph).
        peer_exchange: a parallel.PeerExchange — the graph ends with the fused reduce-scatter + Adam + all-gather
        kernel over NVLink peer memory instead of optimizer.step(): the whole data-parallel step is one graph with
        no NCCL call in it."""
        self.eng, self.opt, self.klw = step_engine, optimizer, float(kl_weight)
        self.split = split_optimizer and peer_exchange is None
        self.sync = grad_sync if peer_exchange is None else None
        self.peer_exchange = peer_exchange
        lst = isinstance(example_x, (list, tuple))
        self.x = [t.clone() for t in example_x] if lst else example_x.clone()
        self.t = [t.clone() for t in example_t] if lst else example_t.clone()
        self.mask = loss_mask.clone() if loss_mask is not None else None
        self.cond = None  # CVAE condition (shock force): a static fp32 (B, cd) buffer the graph reads
        if condition is not None:
            c = condition.float()
            self.cond = (c.unsqueeze(1) if c.dim() == 1 else c).contiguous().clone()
        eng = self.eng
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # warm-up runs allocate every workspace outside the capture.  They must not train: no optimizer
        # step (a graph is re-captured whenever the KL weight changes, i.e. every epoch of the annealing
        # phase, and two stray Adam steps per capture would move the trajectory away from the reference's)
        self.opt._arena()  # moment buffers + device step counter exist before the capture
        self.arena = get_arena(eng.model)
        # ... and must not move model buffers either: BatchNorm running statistics stay untouched
        # (track=False) and the device noise counter is put back, so that a run with N captures (one per
        # annealing epoch) leaves the same state_dict / noise stream as a run with one
        src = eng._noise()
        ctr = getattr(src, "ctr", None)
        ctr_saved = ctr.clone() if ctr is not None else None
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body(False, track=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if ctr_saved is not None and getattr(src, "ctr", None) is not None:
            src.ctr.copy_(ctr_saved)
        elif ctr is None and getattr(src, "ctr", None) is not None:
            src.ctr.zero_()  # the counter was created by the warm-up itself
        eng.always_refresh = True
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs, self.loss = self._body(not self.split)
        self.graph_opt = None
        if self.split:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self.opt.step()
        eng.always_refresh = False

    def _body(self, with_opt, track=True):
        self.opt.zero_grad()
        if self.sync is not None:
            self.sync.begin()
        outputs, loss = self.eng.evaluate(self.x, self.t, self.klw, loss_mask=self.mask, need_grad=True, autograd=False,
                                          condition=self.cond, track=track)
        self.eng.backward()
        if self.sync is not None:
            self.sync.finish()
        if with_opt:
            if self.peer_exchange is not None:
                self.peer_exchange.step()
            else:
                self.opt.step()
        return outputs, loss

    def load(self, x, t, non_blocking=True, mask=None, condition=None):
        if mask is not None and self.mask is not None:
            self.mask.copy_(mask, non_blocking=non_blocking)
        if self.cond is not None:
            if condition is None:
                raise ValueError("this graph was captured for a conditional model: load() needs `condition`")
            self.cond.copy_(condition.reshape(self.cond.shape), non_blocking=non_blocking)
        if isinstance(self.x, list):
            for d, s_ in zip(self.x, x):
                d.copy_(s_, non_blocking=non_blocking)
            for d, s_ in zip(self.t, t):
                d.copy_(s_, non_blocking=non_blocking)
        else:
            self.x.copy_(x, non_blocking=non_blocking)
            self.t.copy_(t, non_blocking=non_blocking)

    def run(self):
        self.graph.replay()
        if not self.split:
            # the replayed optimizer rewrote the parameter arena behind Python's back: the fp16 operand
            # copies packed at the START of this replay are one step old for any eager call that follows
            # (Problem._test_epoch, _sample, the module-level API) -> invalidate the packer token
            self.arena.bump()
        return self.outputs, self.loss

    def apply(self):
        self.graph_opt.replay()
        self.arena.bump()
