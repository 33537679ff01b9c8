"""Data parallelism for the fused step: one process per GPU, gradients summed with NCCL all-reduce
over NVLink 5 / NVSwitch (torch.distributed is the plumbing; the reference has no distributed code
at all — SURVEY.md §5 — so the semantics are defined here and in DESIGN.md §multi-GPU):

  * every rank holds an identical replica and a disjoint shard of the step batch (by rows; by whole
    sequences for dyn_modeling so the roll/fix-up of DynModeling.parse_input stays local);
  * BatchNorm statistics are per rank (local batch), like torch DDP without SyncBN;
  * the flat gradient arena is cut into contiguous buckets, one per sub-network (decoders first:
    their gradients are final first in the fused backward), each all-reduced (SUM) on a side stream as
    soon as the backward reports it ready, overlapping the remaining encoder backward;
  * the fused Adam multiplies by 1/world_size (`grad_prescale`) — sum then scale = average.

Parity definition: N ranks == the oracle run independently on each shard with identical weights,
gradients averaged, one Adam step.
"""
import torch
import torch.distributed as dist


def bucket_ranges(names, offsets, numels, total):
    """Contiguous [start, end) ranges of the flat arena per top-level sub-network, in arena order.
    -> dict prefix -> (start, end).  Padding between parameters stays inside its bucket."""
    order, first = [], {}
    for n in names:
        p = n.split(".")[0]
        if p not in first:
            first[p] = offsets[n]
            order.append(p)
    out = {}
    for i, p in enumerate(order):
        end = first[order[i + 1]] if i + 1 < len(order) else total
        out[p] = (first[p], end)
    return out


def shard_rows(n_rows, world, rank, seq_length=1):
    """Row range of this rank: whole sequences only, remainder spread over the first ranks."""
    assert n_rows % seq_length == 0
    n_seq = n_rows // seq_length
    base, rem = divmod(n_seq, world)
    start = rank * base + min(rank, rem)
    stop = start + base + (1 if rank < rem else 0)
    return start * seq_length, stop * seq_length


def shard_batch(obj, world, rank, seq_length=1):
    """Rows [shard_rows(...)] of every tensor in a parsed batch (tensor / list / tuple / dict, None kept).

    Use it AFTER `parse_input` ran on the whole step batch: DynModeling.parse_input builds its targets
    with a roll over dim 0 (problems.py:785-798), so the last row of a shard needs the first row of the
    next shard (and the pose target of the very last row wraps around to row 0, a quirk of :798 that is
    kept).  Parsing globally on the host and sharding the parsed rows gives every rank exactly the rows
    of the single-process step; parsing each shard locally would change one target row per rank."""
    if obj is None:
        return None
    if isinstance(obj, dict):
        return {k: shard_batch(v, world, rank, seq_length) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(shard_batch(v, world, rank, seq_length) for v in obj)
    a, b = shard_rows(obj.shape[0], world, rank, seq_length)
    return obj[a:b]


class GradSync:
    """Bucketed, overlapped all-reduce of a flat gradient tensor.  Works on any device / backend
    (the CPU + gloo combination is what the host-logic tests use)."""

    def __init__(self, flat_grad, ranges, group=None, overlap=True):
        self.flat, self.ranges, self.group = flat_grad, dict(ranges), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap and flat_grad.is_cuda
        self.stream = torch.cuda.Stream() if self.overlap else None
        self.pending = []
        self.done = set()

    def begin(self):
        self.pending, self.done = [], set()

    def ready(self, prefixes):
        """Gradients of these sub-networks are final: launch their all-reduce."""
        if self.world == 1:
            return
        for p in prefixes:
            if p in self.done or p not in self.ranges:
                continue
            self.done.add(p)
            a, b = self.ranges[p]
            chunk = self.flat[a:b]
            if self.overlap:
                self.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    dist.all_reduce(chunk, group=self.group)
            else:
                self.pending.append(dist.all_reduce(chunk, group=self.group, async_op=True))

    def finish(self):
        """Reduce whatever was not reported and make the result visible to the compute stream."""
        if self.world == 1:
            return
        self.ready([p for p in self.ranges if p not in self.done])
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self.stream)
        for w in self.pending:
            w.wait()
        self.pending = []


class PeerExchange:
    """Gradient exchange fused with the optimizer over NVLink peer memory: reduce-scatter + Adam + all-gather as ONE
    kernel (csrc/peer.cu) instead of an NCCL all-reduce followed by N identical Adam updates.  Rank r owns the r-th
    shard of the flat arena: it sums that shard's gradients over all ranks with peer loads, updates it with its local
    moments, and stores the new parameters into every replica.  No host synchronisation and no NCCL call on the step
    path (the kernel carries its own cross-GPU flags), so the whole data-parallel step is one CUDA graph.

    Set-up (once): every rank exports its gradient arena, parameter arena and a flag block as CUDA IPC handles
    (mmdyn_ipc_export), the handles travel through `torch.distributed` (object all-gather), every rank maps
    its peers' buffers and enables peer access.  One node only (NVLink / NVSwitch)."""

    def __init__(self, arena, optimizer, group=None):
        from . import ops
        if not isinstance(optimizer, __import__("mmdyn_b200.optim", fromlist=["FusedAdam"]).FusedAdam):
            raise TypeError("PeerExchange fuses the exchange with FusedAdam")
        self.arena, self.opt, self.group = arena, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise ValueError("PeerExchange: at most 8 ranks (one NVSwitch box)")
        dev = arena.flat.device
        optimizer._arena()  # moment arenas + device step counter
        self.flags = torch.zeros(2 * self.world, dtype=torch.int32, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        mine = [arena.grad, arena.flat, self.flags]
        ptrs = [[None] * self.world for _ in mine]
        for k, t in enumerate(mine):
            ptrs[k][self.rank] = t.data_ptr()
        if self.world > 1:
            torch.cuda.synchronize(dev)
            with torch.cuda.device(dev):
                exported = [ops.ipc_export(t) for t in mine]
                everyone = [None] * self.world
                dist.all_gather_object(everyone, exported, group=group)
                for p, theirs in enumerate(everyone):
                    if p != self.rank:
                        for k, (handle, off) in enumerate(theirs):
                            ptrs[k][p] = ops.ipc_import(handle, off)  # mapped for THIS device, peer access enabled
            dist.barrier(group=group)  # everybody has mapped everybody before the first kernel may signal
        self.grad_ptrs, self.param_ptrs, self.flag_ptrs = ptrs
        optimizer.grad_prescale = 1.0 / self.world

    def step(self):
        """The optimizer step of problems.py:155 for all replicas at once (call it where optimizer.step() would be)."""
        from . import ops
        opt, arena = self.opt, self.arena
        g = opt.param_groups[0]
        opt._step += 1
        ops.rng_advance(opt._step_dev, 1)
        ops.rng_advance(self.epoch, 1)
        m, v = opt._bufs
        ops.peer_rs_adam_ag(self.grad_ptrs, self.param_ptrs, self.flag_ptrs, m, v, arena.total, self.rank, self.world,
                            g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], opt._step_dev, self.epoch,
                            opt.grad_prescale, opt._flag, self.counter)
        arena.bump()


def attach(step_engine, optimizer, arena, group=None, overlap=True):
    """Wire a StepEngine + fused optimizer for data parallel training; returns the GradSync.
    Usage per step: sync.begin(); loss.backward() (fires bucket hooks); sync.finish(); optimizer.step()."""
    ranges = bucket_ranges(arena.names, arena.offset, arena.numel, arena.total)
    sync = GradSync(arena.grad, ranges, group, overlap)
    step_engine.bucket_hook = sync.ready
    optimizer.grad_prescale = 1.0 / sync.world
    return sync
