"""Micro-benchmark of the BatchNorm streaming kernels at the step's largest shapes (batch 1024)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdyn_b200 import lib, ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev); lib.init(0)
F16, F32 = torch.float16, torch.float32
# (C, rows per group, groups): decoder layers 3 / 2 / 1 / 0 (4 loss-bearing passes), encoder layers 1 / 2 / 3 / 4
shapes = [(32, 1024 * 1024, 4), (64, 256 * 1024, 4), (128, 64 * 1024, 4), (256, 4 * 25 * 1024, 1),
          (32, 1024 * 1024, 1), (64, 256 * 1024, 1), (128, 64 * 1024, 1), (256, 25 * 1024, 1)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, n=10):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for C, rows, G in shapes:
    n = G * rows * C
    x = torch.randn(G * rows, C, device=dev).half(); dy = torch.randn(G * rows, C, device=dev).half(); y = torch.empty_like(x)
    sums = torch.zeros(G, C, 2, device=dev); ab = torch.randn(G, C, 2, device=dev); mi = torch.rand(G, C, 2, device=dev) + 0.5
    sums2 = torch.zeros(G, C, 2, device=dev); coef = torch.zeros(G, C, 4, device=dev)
    t1 = timeit(lambda: ops.bn_stats(x, sums, G, rows, C))
    t2 = timeit(lambda: ops.bn_swish_fwd(x, ab, y, G, rows, C))
    t3 = timeit(lambda: ops.bn_swish_bwd_reduce(x, ab, mi, dy, sums2, G, rows, C))
    t4 = timeit(lambda: ops.bn_bwd_apply(x, ab, mi, sums2, dy, None, None, coef, G, rows, C, 1.0))
    gb = n * 2 / 1e9
    print(f"C={C:4d} rows={rows:8d} G={G}  stats {t1*1e3:7.1f}us {gb/t1*1e3:6.0f} GB/s | fwd {t2*1e3:7.1f}us {2*gb/t2*1e3:6.0f} | "
          f"reduce {t3*1e3:7.1f}us {2*gb/t3*1e3:6.0f} | apply(+coef) {t4*1e3:7.1f}us {3*gb/t4*1e3:6.0f}", flush=True)
