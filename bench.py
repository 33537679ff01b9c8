#!/usr/bin/env python
"""Benchmark of the cnn-mvae (visuotactile + pose) training step — BASELINE.json's metric:
"cnn-mvae train samples/s at 1/2/4/8 B200 + roofline %; CPU-host baseline".

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

One step = one training iteration of `SeqModeling._train_epoch` (problems.py:150-155) on one
synthetic, dataset-shaped batch: zero_grad, the 7 sub-sampled MVAE passes, backward, Adam.
One sample = one (visual, tactile, pose) triple going through all 7 passes.

Prints ONE JSON line (rank 0):
  value      samples/s, whole job, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public host API: every step copies that step's pinned host
             batch to the device and reads the loss back
  roofline   the dominant kernel of the step, timed live with CUDA events in a separate pass
  cpu_baseline  the oracle (CPU port of the reference step) on the host cores, bounded sample
Weak scaling: the per-GPU batch is fixed; gradients are summed with NCCL all-reduce.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "cnn-mvae train samples/s"
UNIT = "samples/s"
FLOP_PER_SAMPLE = 2021.24e6  # necessary fwd + 2x bwd work of cnn-mvae + pose
KW = dict(condition_dim=0, input_dim=4096, architecture="cnn", conditional=False, categorical_conditions=False,
          latent_size=256, use_pose=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML in-process (a query is
    microseconds, so even a 0.2 s region yields samples), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            try:  # CUDA_VISIBLE_DEVICES may renumber the devices: go through the PCI address
                bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = self.handle = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(h) / 1e3
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(h))
        act = lambda bit: "Active" if r & bit else "Not Active"
        # NVML reason bits: 0x8 hw_slowdown, 0x40 hw_thermal_slowdown, 0x20 sw_thermal_slowdown, 0x4 sw_power_cap
        return [str(sm), str(self.max_sm), str(pw), act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.02)
                    continue
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                self.nvml = None  # NVML failed mid-run: fall back to nvidia-smi
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(float(s[2]) for s in self.samples)}


def synth_batch(B, seed, device=None, pin=False):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    t = [r(B, 3, 64, 64), r(B, 3, 64, 64), r(B, 7), r(B, 3, 64, 64), r(B, 3, 64, 64), r(B, 7)]
    if pin:
        t = [x.pin_memory() for x in t]
    if device is not None:
        t = [x.to(device) for x in t]
    return t[:3], t[3:]


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (baseline/_ref, copied from /root/reference by
# __graft_entry__.build()) through its own SeqModeling._evaluate_model + torch.optim.Adam on the host cores;
# the oracle port only if that copy is absent
# ---------------------------------------------------------------------------------------------
REF_SAMPLE_B = 128  # the reference's own default --batchsize (main.py:25) and its best CPU throughput


def cpu_steps(sample_B, steps, warmup, threads):
    """-> (seconds per step, kind, description)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    if ref_runner.available():
        times, _, root = ref_runner.time_step(sample_B, steps, warmup, "cpu", threads)
        where = "baseline/_ref" if root.endswith("_ref") else root
        return times, "reference", f"the reference's own mmdyn.pytorch step (SeqModeling._evaluate_model + backward + torch.optim.Adam, {where})"
    from oracle import mmdyn_oracle as orc
    from mmdyn_b200.pytorch.models.models import setup_model
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = setup_model("cnn-mvae", cross_modal=True, **KW)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    x, t = synth_batch(sample_B, 1)
    st = {"step": 0, "m": [torch.zeros_like(sd[k]) for k in pkeys], "v": [torch.zeros_like(sd[k]) for k in pkeys]}
    times = []
    for i in range(warmup + steps):
        noises = [orc.draw_pass_noise(sample_B, hv, ht) for (hv, ht, hp) in orc.MVAE_PASSES_POSE]
        t0 = time.perf_counter()
        orc.train_step(sd, pkeys, "mvae+pose", {"x": x, "targets": t}, 1.0 / 50, 1000.0, noises, st)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, "port", "oracle port of the reference step (baseline/_ref absent)"


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_B = min(args.batch, REF_SAMPLE_B)
    # ~0.4-1.3 s per 128-sample step: cap the timed steps so the arm ends within ~2 minutes
    n_steps, n_warm = min(args.steps, 40), min(max(args.warmup, 1), 2)
    times, kind, what = cpu_steps(sample_B, n_steps, n_warm, threads)
    tot = sum(times)
    v = sample_B * len(times) / tot
    sample = (f"{len(times)} timed steps x {sample_B} samples (each step = a {sample_B}-sample SAMPLE of the per-GPU batch "
              f"{args.batch} the B200 arm runs; CPU throughput is flat-to-falling in batch, BASELINE.md §2): {what}; fp32, {threads} threads")
    cfg = workload_config(args, args.batch)
    cfg["reference_sample_batch"] = sample_B
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": n_warm, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def reference_cuda(B, dev, steps=8, warmup=3):
    """Same-box comparator BASELINE.md §3 / SURVEY §0 name: the reference's own modules `.to('cuda')`, stock
    PyTorch eager (cuDNN / cuBLAS), same step, same batch.  Informational (`extra.reference_cuda`)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    if not ref_runner.available():
        return {"unavailable": "baseline/_ref absent"}
    try:
        times, lv, _ = ref_runner.time_step(B, steps, warmup, dev)
        ms = 1e3 * sum(times) / len(times)
        return {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": B, "steps": len(times),
                "what": "unmodified reference modules .to('cuda'), stock PyTorch eager fp32 (TF32 off), 1 GPU, "
                        "inputs resident, loss.item() per step as problems.py:156", "final_loss": lv}
    except Exception as e:  # noqa: BLE001 - informational leg
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    finally:
        torch.cuda.empty_cache()


def synth_batch_u8(B, seed):
    """The same batch in the dataset's native precision: 8-bit frames (the reference reads PNG renders and
    divides by 255, datasets.py:23-31), HWC uint8, pinned; poses fp32.  Order as synth_batch: v, t, p, tv, tt, tp."""
    g = torch.Generator().manual_seed(seed)
    img = lambda: torch.randint(0, 256, (B, 64, 64, 3), generator=g, dtype=torch.uint8).pin_memory()
    vec = lambda: torch.rand(B, 7, generator=g).pin_memory()
    return [img(), img(), vec(), img(), img(), vec()]


def dyn_shard(rank, world, S, L, seed):
    """configs[3]: this rank's rows of a dyn_modeling step batch of world*S sequences x L frames, parsed as
    DynModeling.parse_input parses the WHOLE batch (targets = roll(-1) over the frame axis, every
    last-of-sequence image row <- the resting-state target, pose target = bare roll incl. the wrap-around of the
    very last row to row 0, problems.py:765-803).  A shard's last row needs the next shard's first row, so the
    rank builds its S sequences plus the first sequence of the next rank (cyclically) and parses those; the
    rows of its own sequences are then exactly the rows of the global parse (tests/test_host_cpu.py)."""
    def seq(i):
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        r = lambda *s_: torch.rand(*s_, generator=g)
        return r(L, 3, 64, 64), r(L, 3, 64, 64), r(L, 7), r(1, 3, 64, 64), r(1, 3, 64, 64)
    ids = [rank * S + j for j in range(S)] + [((rank + 1) % world) * S]
    parts = [seq(i) for i in ids]
    vis, tac, pose = (torch.cat([p[k] for p in parts]) for k in range(3))
    rest_v = torch.cat([p[3].expand(L, -1, -1, -1) for p in parts])
    rest_t = torch.cat([p[4].expand(L, -1, -1, -1) for p in parts])

    def shifted(x, rest):
        t = torch.roll(x, -1, dims=0)
        t[L - 1::L] = rest[L - 1::L]
        return t
    n = S * L
    x = [vis[:n], tac[:n], pose[:n]]
    t = [shifted(vis, rest_v)[:n], shifted(tac, rest_t)[:n], torch.roll(pose, -1, dims=0)[:n]]
    return x, t


def workload_config(args, B):
    if getattr(args, "problem", "seq") == "dyn":
        wl = ("cnn-mvae --input-type visuotactile --use-pose dyn_modeling with missing-modality sub-sampling: all S*L frames "
              f"of {B // args.seq_len} sequences x {args.seq_len} frames per GPU, one-step targets by roll/fix-up, 7 passes + "
              "backward + Adam (BASELINE.json configs[3]), sharded by whole sequences")
    else:
        wl = ("cnn-mvae --input-type visuotactile --use-pose seq_modeling: 7 sub-sampled passes + backward + Adam "
              f"(BASELINE.json configs[2]/[4]), per-GPU batch {B}")
    return {"workload": wl,
            "per_gpu_batch": B, "global_batch": B * args.gpus, "latent": 256, "image": "3x64x64 x2 + pose 7",
            "parallelism": f"dp{args.gpus}" if args.gpus > 1 else "single",
            "l2": "per-step working set (activations + inputs, ~5 MB/sample) exceeds the 126 MB L2"}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from mmdyn_b200 import engine, lib, noise, ops, optim
    from mmdyn_b200.pytorch.models.models import setup_model

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # captured NCCL collectives: the watchdog's event queries must not run into the capture
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        # NCCL prints its version banner on stdout when the first communicator comes up: point fd 1 at
        # stderr while that happens, so that stdout carries the JSON line and nothing else
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dyn = args.problem == "dyn"
    if dyn:
        args.batch = args.sequences * args.seq_len  # rows entering the model per GPU and step
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    pk = peaks()

    torch.manual_seed(0)  # identical replicas
    model = setup_model("cnn-mvae", cross_modal=True, **KW).to(dev)
    eng = engine.StepEngine(model, "mvae", use_pose=True, pose_multiplier=1000.0,
                            noise_src=noise.DeviceNoise(seed=1234 + rank))
    opt = optim.FusedAdam(model, lr=1e-3)
    opt.grad_prescale = 1.0 / world
    klw = 1.0 / 50
    if dyn:
        dev_batches = []
        for i in range(3):
            xb, tb = dyn_shard(rank, world, args.sequences, args.seq_len, 10 + i)
            dev_batches.append(([a.to(dev) for a in xb], [a.to(dev) for a in tb]))
        del xb, tb
    else:
        dev_batches = [synth_batch(B, 10 + 3 * rank + i, device=dev) for i in range(3)]
    # e2e feeds the step from pinned HOST memory in the dataset's native precision: uint8 frames (converted to
    # the fp32 NCHW tensors the model reads by mmdyn_frames_u8_to_f32) + fp32 poses
    host_batches = [synth_batch_u8(B, 10 + 3 * rank + i) for i in range(3)]
    u8_table = ops.resize_table(64, 64, 64, 64).to(dev)

    arena = engine.get_arena(model, dev)
    from mmdyn_b200 import parallel
    # eager path: bucketed all-reduce launched from the backward as sub-networks finish (overlap);
    # graphed path: one all-reduce of the flat arena between the backward graph and the Adam graph
    bucketed = parallel.attach(eng, opt, arena, overlap=True) if (world > 1 and args.no_graph) else None

    def sync_grads():
        if bucketed is not None:
            bucketed.finish()
            bucketed.begin()
        elif world > 1:
            dist.all_reduce(arena.grad)

    # one eager step: builds workspaces, counts this library's launches per step
    c0 = lib.launch_count()
    opt.zero_grad()
    _, loss = eng.evaluate(dev_batches[0][0], dev_batches[0][1], klw, need_grad=True, autograd=False)
    eng.backward()
    sync_grads()
    opt.step()
    torch.cuda.synchronize()
    launches_per_step = lib.launch_count() - c0

    use_graph = not args.no_graph
    gstep = None
    dp_mode = "single" if world == 1 else ("eager: bucketed all-reduce overlapped with backward" if args.no_graph else "")
    px = None
    if use_graph:
        if world > 1 and args.dp == "peer":
            # data parallel as ONE CUDA graph with no NCCL call in it: the graph ends with the fused
            # reduce-scatter + Adam + all-gather kernel over NVLink peer memory (csrc/peer.cu): rank r sums the r-th
            # shard of all ranks' gradient arenas with peer loads, updates it, and stores the new parameters into
            # every replica; cross-GPU flags inside the kernel replace the host-side collective
            try:
                px = parallel.PeerExchange(arena, opt)
                gstep = engine.GraphedTrainStep(eng, opt, dev_batches[0][0], dev_batches[0][1], klw, peer_exchange=px)
                dp_mode = ("one CUDA graph: fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory "
                           "(CUDA IPC mapped arenas, in-kernel flags, no NCCL on the step path)")
            except Exception as e:  # noqa: BLE001 - e.g. no peer access: fall back to NCCL
                if rank == 0:
                    print(f"[bench] peer exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
                px, gstep = None, None
        if world > 1 and gstep is None and args.dp != "split":
            # the gradient arena is cut into one bucket per sub-network, each bucket's NCCL all-reduce is launched
            # from the backward on a side stream as soon as that sub-network's gradients are final (decoders first)
            # and overlaps the encoder backward; the fused Adam (grad_prescale = 1/N) closes the graph
            try:
                sync_in_graph = parallel.attach(eng, opt, arena, overlap=True)
                gstep = engine.GraphedTrainStep(eng, opt, dev_batches[0][0], dev_batches[0][1], klw,
                                                grad_sync=sync_in_graph)
                dp_mode = "one CUDA graph: bucketed NCCL all-reduce captured, overlapped with backward, fused Adam"
            except Exception as e:  # noqa: BLE001 - fall back to the split-graph scheme below
                if rank == 0:
                    print(f"[bench] NCCL capture failed ({type(e).__name__}: {e}); using split graphs", file=sys.stderr)
                eng.bucket_hook = None
                gstep = None
        if gstep is None:
            gstep = engine.GraphedTrainStep(eng, opt, dev_batches[0][0], dev_batches[0][1], klw, split_optimizer=world > 1)
            if world > 1:
                dp_mode = "backward graph + flat all-reduce + optimizer graph"
    in_graph_sync = gstep is not None and (gstep.sync is not None or px is not None)  # the graph holds the exchange

    def step_resident(i):
        x, t = dev_batches[i % 3]
        if gstep is not None:
            gstep.run()  # inputs already resident in the graph's static device buffers
            if world > 1 and not in_graph_sync:
                sync_grads()
                gstep.apply()
            return gstep.loss
        opt.zero_grad()
        _, l = eng.evaluate(x, t, klw, need_grad=True, autograd=False, want_outputs=False)
        eng.backward()
        sync_grads()
        opt.step()
        return l

    # e2e: every step's batch comes from pinned HOST memory and its loss goes back to the host, all inside
    # the timed region.  Two sets of static device inputs (and, in graph mode, two captures of the same
    # step reading one set each) alternate: the host->device copy of step i+1 lands directly in the set
    # step i does not use, on a copy stream, while step i computes (no staging copy); the loss of step i
    # is copied to a pinned slot asynchronously and read on the host while step i+1 runs.
    copy_stream = torch.cuda.Stream()
    gsteps = [gstep, None]
    if gstep is not None:
        gsteps[1] = engine.GraphedTrainStep(eng, opt, dev_batches[1][0], dev_batches[1][1], klw,
                                            split_optimizer=world > 1 and not in_graph_sync, grad_sync=gstep.sync,
                                            peer_exchange=px)
        in_sets = [g_.x + g_.t for g_ in gsteps]
    else:
        in_sets = [[torch.empty_like(a) for a in dev_batches[0][0] + dev_batches[0][1]] for _ in range(2)]
    u8_sets = [[torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host_batches[0]] for _ in range(2)]
    staged_evt = [torch.cuda.Event(), torch.cuda.Event()]
    consumed_evt = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"primed": False, "pending": None, "last": float("nan")}

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed_evt[slot])  # the step that last read this set has finished
            for dst, stg, src in zip(in_sets[slot], u8_sets[slot], host_batches[i % 3]):
                stg.copy_(src, non_blocking=True)           # host -> device, uint8 frames / fp32 poses
                if src.dtype == torch.uint8:
                    ops.frames_u8_to_f32(stg, None, u8_table, dst)  # /255 -> fp32 NCHW, on the copy stream
                else:
                    dst.copy_(stg, non_blocking=True)
            staged_evt[slot].record(copy_stream)

    def step_e2e(i):
        if not e2e_state["primed"]:
            for ev in consumed_evt:
                ev.record()
            stage(i)
            e2e_state["primed"] = True
        slot = i % 2
        stage(i + 1)  # prefetch the next batch while this step runs
        cur = torch.cuda.current_stream()
        cur.wait_event(staged_evt[slot])
        if gstep is not None:
            g_ = gsteps[slot]
            g_.run()
            if world > 1 and not in_graph_sync:
                sync_grads()
                g_.apply()
            l = g_.loss
        else:
            xs, ts = in_sets[slot][:3], in_sets[slot][3:]
            opt.zero_grad()
            _, l = eng.evaluate(xs, ts, klw, need_grad=True, autograd=False, want_outputs=False)
            eng.backward()
            sync_grads()
            opt.step()
        consumed_evt[slot].record(cur)
        loss_host[slot].copy_(l.reshape(1), non_blocking=True)  # the step's result goes back to the host ...
        loss_evt[slot].record(cur)
        if e2e_state["pending"] is not None:                     # ... and is read one step later, off the critical path
            ps = e2e_state["pending"]
            loss_evt[ps].synchronize()
            e2e_state["last"] = float(loss_host[ps][0])
        e2e_state["pending"] = slot
        return e2e_state["last"]

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(W):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, K)
    for i in range(2):
        step_e2e(i)
    e2e_off = 2
    ms_e2e = timed(lambda i: step_e2e(i + e2e_off), K)
    sampler.stop_flag = True
    final_loss = float(step_resident(0).item())
    assert final_loss == final_loss and final_loss < 1e9, f"training diverged: loss={final_loss}"

    value = world * B * K / (ms_total / 1e3)
    e2e_v = world * B * K / (ms_e2e / 1e3)
    h2d = sum(a.numel() * a.element_size() for a in host_batches[0])

    # ---- sustained: the same resident step for >= 5 s (power-capped steady state, not a burst) ----
    sustained = None
    if not args.no_sustained:
        n_s = max(K, int(5.5e3 / (ms_total / K)))
        ms_s = timed(step_resident, n_s)
        sustained = {"value": world * B * n_s / (ms_s / 1e3), "unit": UNIT, "steps": n_s, "seconds": ms_s / 1e3,
                     "ms_per_step": ms_s / n_s}

    # ---- roofline pass: every launch of this library timed with CUDA events (eager, untimed run) ----
    roof, roof_hbm, table, families, tensor_all = None, None, [], [], None
    if rank == 0:
        eng.set_concurrent(False)  # serial branches: per-kernel events must not time-share the SMs
        saved_hook, eng.bucket_hook = eng.bucket_hook, None  # rank-0-only pass: no collectives
        for _ in range(2):
            ops.start_profile()
            opt.zero_grad()
            eng.evaluate(dev_batches[0][0], dev_batches[0][1], klw, need_grad=True, autograd=False, want_outputs=False)
            eng.backward()
            opt.step()
            prof = ops.stop_profile()
        eng.bucket_hook = saved_hook
        tot_ms = sum(d["ms"] for d in prof.values())
        for tag, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            table.append({"kernel": tag, "launches": d["count"], "ms": round(d["ms"], 4),
                          "share": round(d["ms"] / tot_ms, 4),
                          "tflops": round(d["flops"] / d["ms"] / 1e9, 2) if d["flops"] else None,
                          "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["bytes"] else None})
        # The profile tags are per LAYER for the GEMMs ("deconv3.fwd") and per KERNEL for the streaming ops; the
        # roofline is reported per kernel NAME, so the layer tags are pooled into the kernel that runs them.
        fam = {}
        for tag, d in prof.items():
            name = ops.kernel_family(tag)
            f = fam.setdefault(name, dict(count=0, ms=0.0, flops=0.0, bytes=0.0, tags=[]))
            for k_ in ("count", "ms", "flops", "bytes"):
                f[k_] += d[k_]
            f["tags"].append(tag)
        try:  # DRAM traffic per launch from the ncu --set full capture of THIS build (same batch only)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            tr = tr if tr.get("batch") == B else {}
        except Exception:
            tr = {}

        def block(name, f):
            t_tensor = f["flops"] / (pk["tf_sust"] * 1e12)
            t_hbm = f["bytes"] / (pk["hbm"] * 1e9)
            common = {"kernel": name, "layers": sorted(f["tags"]) if len(f["tags"]) > 1 else None,
                      "traffic": (tr.get(name) or {}).get("dram_bytes_per_launch"),
                      "traffic_source": (tr.get(name) or {}).get("source"),
                      "launch_ms": f["ms"] / f["count"], "launches": f["count"], "share_of_step": f["ms"] / tot_ms,
                      "algorithmic_bytes_per_launch": f["bytes"] / f["count"],
                      "algorithmic_flops_per_launch": f["flops"] / f["count"]}
            if t_tensor > t_hbm:
                ach = f["flops"] / (f["ms"] / 1e3) / 1e12
                return dict(common, bound="tensor", achieved=ach, peak=pk["tf_sust"], unit="TFLOP/s", frac=ach / pk["tf_sust"],
                            peak_source=pk["src"] + ", sustained bf16/fp16 GEMM (kernel timed inside a long step)")
            ach = f["bytes"] / (f["ms"] / 1e3) / 1e9
            return dict(common, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                        peak_source=pk["src"])
        ranked = sorted(fam.items(), key=lambda kv: -kv[1]["ms"])
        blocks = [block(n_, f_) for n_, f_ in ranked if f_["flops"] or f_["bytes"]]
        roof = blocks[0]                                                   # the dominant kernel of the step
        roof_hbm = next((b_ for b_ in blocks if b_["bound"] == "hbm"), None)  # and the top HBM-bound one
        families = [{"kernel": b_["kernel"], "bound": b_["bound"], "ms": round(b_["launch_ms"] * b_["launches"], 4),
                     "share": round(b_["share_of_step"], 4), "frac": round(b_["frac"], 4)} for b_ in blocks[:8]]
        # all tensor-core layers together (north_star's FLOP-weighted figure)
        tc = [f_ for n_, f_ in fam.items() if n_.startswith(("igemm_tma", "wgrad_tma", "conv1_"))]
        tc_ms, tc_fl = sum(f_["ms"] for f_ in tc), sum(f_["flops"] for f_ in tc)
        tensor_all = {"ms": tc_ms, "tflops": tc_fl / tc_ms / 1e9, "frac_of_sustained_peak": tc_fl / tc_ms / 1e9 / pk["tf_sust"]} if tc_ms else None
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"batch": B, "ms_per_step_events_sum": tot_ms, "kernels": table, "families": families,
                           "tensor_layers": tensor_all}, f, indent=1)

    # ---- config 5 grid (BASELINE.json configs[4]): GLOBAL batch 64..4096 on THIS run's N GPUs; 128 = --batchsize default ----
    sweep = None
    if world > 1:
        dist.barrier()  # rank 0 ran the roofline pass alone: meet before the ranks replay exchange kernels again
    if use_graph and not args.no_sweep and not dyn:
        eng.set_concurrent(True)
        sweep = []
        for Bg in (64, 128, 256, 512, 1024, 2048, 4096):
            Bs = Bg // world
            if Bs < 8:
                continue
            xb, tb = synth_batch(Bs, 99 + rank, device=dev)
            g2 = engine.GraphedTrainStep(eng, opt, xb, tb, klw, split_optimizer=world > 1 and not in_graph_sync,
                                         grad_sync=gstep.sync if in_graph_sync else None, peer_exchange=px)

            def one():
                g2.run()
                if world > 1 and not in_graph_sync:
                    dist.all_reduce(arena.grad)
                    g2.apply()
            for _ in range(5):
                one()
            n = 30
            ms = timed(lambda i: one(), n) / n
            sweep.append({"global_batch": Bg, "per_gpu_batch": Bs, "n_gpus": world, "ms_per_step": round(ms, 4),
                          "samples_per_s": round(Bg / ms * 1e3, 1)})
            del g2, xb, tb

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_ref_cuda and not dyn:
        ref_cuda = reference_cuda(B, dev)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sb = min(B, REF_SAMPLE_B)
        ts, kind, what = cpu_steps(sb, 6, 1, threads)
        cpu = {"value": sb * len(ts) / sum(ts), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{len(ts)} steps x {sb} samples of the same step: {what}; fp32, {threads} threads"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands (10-bit mantissa, TF32-equivalent) / f32 accumulate",
            "data": "synthetic", "config": workload_config(args, B),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K,
                    "input_format": "pinned host uint8 HWC frames (the dataset's native 8-bit renders) + fp32 poses; "
                                    "uint8 -> fp32 NCHW /255 on the device (mmdyn_frames_u8_to_f32) inside the timed region"},
            "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step),
            "cuda_graph": bool(use_graph), "data_parallel_mode": dp_mode, "clocks": sampler.summary(), "roofline": roof, "roofline_hbm": roof_hbm,
            "cpu_baseline": cpu,
            "extra": {"sustained": sustained, "reference_cuda": ref_cuda, "grid": sweep, "kernel_families": families,
                      "tensor_layers_all": tensor_all},
            "step_tflops": value * FLOP_PER_SAMPLE / 1e12 / world,
            "step_frac_of_tensor_peak": value * FLOP_PER_SAMPLE / 1e12 / world / pk["tf_sust"],
            "final_loss": final_loss, "top_kernels": table[:6],
        }
        print(json.dumps(out))
        sys.stdout.flush()
    if world > 1:
        # CUDA graphs that captured NCCL kernels are still alive here; tearing the communicator down under them
        # can hang (measured: the JSON line was out, destroy_process_group never returned).  Drain, meet, leave.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MMDYN_BENCH_BATCH", 1024)),
                    help="per-GPU batch (weak scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl", "split"],
                    help="data-parallel exchange of the graphed step: peer = fused reduce-scatter + Adam + all-gather kernel "
                         "over NVLink peer memory (default); nccl = bucketed NCCL all-reduces captured in the graph, "
                         "overlapped with backward; split = backward graph + flat all-reduce + optimizer graph")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip extra.reference_cuda (stock eager reference on this GPU)")
    ap.add_argument("--no-sustained", action="store_true", help="skip extra.sustained (>= 5 s of the resident step)")
    ap.add_argument("--problem", default="seq", choices=["seq", "dyn"],
                    help="seq: BASELINE configs[2] (default); dyn: configs[3], dyn_modeling over --sequences x --seq-len frames per GPU")
    ap.add_argument("--sequences", type=int, default=128, help="dyn: sequences per GPU and step (--batchsize default 128)")
    ap.add_argument("--seq-len", type=int, default=50, help="dyn: frames per sequence (50 in exp 1/2)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-GPU batch sweep (64-4096, 30 graph replays each)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table here (JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
