"""Debug helper for parallel.PeerExchange: (1) the kernel alone with world = 1 against FusedAdam; under torchrun
(2) the CUDA-IPC mapping: read a peer's tensor through the mapped pointer with a torch copy, then the kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mmdyn_b200 import ops


def single():
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    n = 1 << 20
    g = torch.Generator(device=dev).manual_seed(1)
    p0 = torch.randn(n, device=dev, generator=g)
    gr = torch.randn(n, device=dev, generator=g)
    pa, pb = p0.clone(), p0.clone()
    ma, va, mb, vb = (torch.zeros(n, device=dev) for _ in range(4))
    step = torch.ones(1, dtype=torch.int64, device=dev)
    epoch = torch.ones(1, dtype=torch.int64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    flags = torch.zeros(2, dtype=torch.int32, device=dev)
    ops.adam_flat_devstep(pa, gr, ma, va, n, 1e-3, 0.9, 0.999, 1e-8, 0.0, step, 0.5, flag=flag)
    ops.peer_rs_adam_ag([gr.data_ptr()], [pb.data_ptr()], [flags.data_ptr()], mb, vb, n, 0, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0,
                        step, epoch, 0.5, flag, cnt)
    torch.cuda.synchronize()
    print("world=1 kernel vs adam_flat: max abs diff", (pa - pb).abs().max().item(), (ma - mb).abs().max().item(),
          "flags", flags.tolist(), "counter", cnt.item(), flush=True)


def multi():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    t = torch.full((1 << 16,), float(rank + 1), device=dev)
    torch.cuda.synchronize()
    allh = [None] * world
    dist.all_gather_object(allh, ops.ipc_export(t))

    class _P:  # minimal tensor-like wrapper so that ops._ptr accepts a raw pointer
        is_cuda = True

        def __init__(self, ptr):
            self._p = ptr

        def data_ptr(self):
            return self._p
    for p in range(world):
        if p == rank:
            continue
        handle, off = allh[p]
        ptr = ops.ipc_import(handle, off)
        print(rank, "peer", p, "offset", off, "mapped at", hex(ptr), flush=True)
        dst = torch.zeros(1 << 16, dtype=torch.float16, device=dev)
        ops.f32_to_f16(_P(ptr), dst, 1 << 16, 1.0)
        torch.cuda.synchronize()
        print(rank, "read through own kernel:", dst[:3].tolist(), flush=True)
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        multi()
    else:
        single()
