import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdyn_b200.pytorch.utils.datasets import DeviceFrameStore
dev = "cuda"
n = 4096
frames = torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8, device=dev)
st = DeviceFrameStore(frames, (64, 64), dev)
out = torch.empty(n, 3, 64, 64, device=dev)
idx = torch.randperm(n, device=dev)
for _ in range(3): st.images(idx, out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): st.images(idx, out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
byt = n * (256 * 256 * 3 + 3 * 64 * 64 * 4)
print(f"frames_u8_to_f32: {n} frames 256x256 -> 64x64 in {ms:.3f} ms = {n/ms*1e3:.0f} frames/s, {byt/ms/1e6:.0f} GB/s algorithmic")
