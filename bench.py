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
# reference arm / cpu baseline: the oracle port of the reference step on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_steps(sample_B, steps, warmup, threads):
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
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_B = min(args.batch, 128)
    # ~0.4 s per 128-sample step on 16 cores: cap the timed steps so the arm ends within ~2 minutes
    times = cpu_steps(sample_B, min(args.steps, 40), min(args.warmup, 2), threads)
    tot = sum(times)
    v = sample_B * len(times) / tot
    sample = f"{len(times)} steps x {sample_B} samples of the same 7-pass cnn-mvae+pose step (fp32, torch CPU ops)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, B):
    return {"workload": "cnn-mvae --input-type visuotactile --use-pose seq_modeling: 7 sub-sampled passes + backward + Adam "
                        f"(BASELINE.json configs[2]/[4]), per-GPU batch {B}",
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
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    pk = peaks()

    torch.manual_seed(0)  # identical replicas
    model = setup_model("cnn-mvae", cross_modal=True, **KW).to(dev)
    eng = engine.StepEngine(model, "mvae", use_pose=True, pose_multiplier=1000.0,
                            noise_src=noise.DeviceNoise(seed=1234 + rank))
    opt = optim.FusedAdam(model, lr=1e-3)
    opt.grad_prescale = 1.0 / world
    klw = 1.0 / 50
    dev_batches = [synth_batch(B, 10 + 3 * rank + i, device=dev) for i in range(3)]
    host_batches = [synth_batch(B, 10 + 3 * rank + i, pin=True) for i in range(3)]

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
    if use_graph:
        if world > 1 and args.nccl_in_graph:
            # experimental: capture the bucketed NCCL all-reduces inside the step graph.  On this stack
            # (torch 2.11 / NCCL 2.28.9) the capture HANGS (measured in round 1, 2 GPUs), so the
            # default for data parallelism is the split scheme below.
            try:
                sync_in_graph = parallel.attach(eng, opt, arena, overlap=True)
                gstep = engine.GraphedTrainStep(eng, opt, dev_batches[0][0], dev_batches[0][1], klw,
                                                grad_sync=sync_in_graph)
                dp_mode = "one graph: bucketed all-reduce captured, overlapped with backward"
            except Exception as e:  # noqa: BLE001 - fall back to the split-graph scheme below
                if rank == 0:
                    print(f"[bench] NCCL capture failed ({type(e).__name__}: {e}); using split graphs", file=sys.stderr)
                eng.bucket_hook = None
                gstep = None
        if gstep is None:
            gstep = engine.GraphedTrainStep(eng, opt, dev_batches[0][0], dev_batches[0][1], klw, split_optimizer=world > 1)
            if world > 1:
                dp_mode = "backward graph + flat all-reduce + optimizer graph"
    in_graph_sync = gstep is not None and gstep.sync is not None

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
                                            split_optimizer=world > 1, grad_sync=gstep.sync)
        in_sets = [g_.x + g_.t for g_ in gsteps]
    else:
        in_sets = [[torch.empty_like(a, device=dev) for a in host_batches[0][0] + host_batches[0][1]] for _ in range(2)]
    staged_evt = [torch.cuda.Event(), torch.cuda.Event()]
    consumed_evt = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"primed": False, "pending": None, "last": float("nan")}

    def stage(i):
        x, t = host_batches[i % 3]
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed_evt[slot])  # the step that last read this set has finished
            for dst, src in zip(in_sets[slot], x + t):
                dst.copy_(src, non_blocking=True)
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
    h2d = sum(a.numel() * a.element_size() for a in host_batches[0][0] + host_batches[0][1])

    # ---- roofline pass: every launch of this library timed with CUDA events (eager, untimed run) ----
    roof, table = None, []
    if rank == 0:
        eng.set_concurrent(False)  # serial branches: per-kernel events must not time-share the SMs
        eng.bucket_hook = None     # rank-0-only pass: no collectives
        for _ in range(2):
            ops.start_profile()
            opt.zero_grad()
            eng.evaluate(dev_batches[0][0], dev_batches[0][1], klw, need_grad=True, autograd=False, want_outputs=False)
            eng.backward()
            opt.step()
            prof = ops.stop_profile()
        tot_ms = sum(d["ms"] for d in prof.values())
        for tag, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            table.append({"kernel": tag, "launches": d["count"], "ms": round(d["ms"], 4),
                          "share": round(d["ms"] / tot_ms, 4),
                          "tflops": round(d["flops"] / d["ms"] / 1e9, 2) if d["flops"] else None,
                          "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["bytes"] else None})
        top_tag, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        # which roof binds the dominant kernel: time at the tensor peak vs time at the HBM peak for its
        # ALGORITHMIC flops / bytes (e.g. the 3-channel logits layer is a GEMM but HBM-bound: 56 FLOP/B)
        t_tensor = top["flops"] / (pk["tf_sust"] * 1e12)
        t_hbm = top["bytes"] / (pk["hbm"] * 1e9)
        common = {"kernel": top_tag, "traffic": None, "launch_ms": top["ms"] / top["count"],
                  "launches": top["count"], "share_of_step": top["ms"] / tot_ms,
                  "algorithmic_bytes_per_launch": top["bytes"] / top["count"],
                  "algorithmic_flops_per_launch": top["flops"] / top["count"]}
        if t_tensor > t_hbm:
            ach = top["flops"] / (top["ms"] / 1e3) / 1e12
            roof = dict(common, bound="tensor", achieved=ach, peak=pk["tf_sust"], unit="TFLOP/s", frac=ach / pk["tf_sust"],
                        peak_source=pk["src"] + ", sustained bf16/fp16 GEMM")
        else:
            ach = top["bytes"] / (top["ms"] / 1e3) / 1e9
            roof = dict(common, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                        peak_source=pk["src"])
        try:  # DRAM traffic of the dominant kernel from the committed ncu capture (same batch only)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if tr.get("batch") == B and top_tag in tr:
                roof["traffic"] = tr[top_tag]["traffic_bytes_largest_launch"]
                roof["traffic_note"] = ("largest launch of this kernel, ncu --set full; algorithmic bytes of that launch: %d"
                                        % tr[top_tag]["algorithmic_bytes_largest_launch"])
        except Exception:
            pass
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"batch": B, "ms_per_step_events_sum": tot_ms, "kernels": table}, f, indent=1)

    # ---- batch sweep (BASELINE.json configs[4]: batch 64-4096; 128 = the reference's --batchsize default) ----
    sweep = None
    if rank == 0 and world == 1 and use_graph and not args.no_sweep:
        eng.set_concurrent(True)
        sweep = []
        for Bs in (64, 128, 256, 512, 1024, 2048, 4096):
            if Bs == B:
                sweep.append({"per_gpu_batch": Bs, "ms_per_step": ms_total / K, "samples_per_s": value})
                continue
            xb, tb = synth_batch(Bs, 99, device=dev)
            g2 = engine.GraphedTrainStep(eng, opt, xb, tb, klw)
            for _ in range(5):
                g2.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 30
            e0.record()
            for _ in range(n):
                g2.run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            sweep.append({"per_gpu_batch": Bs, "ms_per_step": round(ms, 4), "samples_per_s": round(Bs / ms * 1e3, 1)})
            del g2, xb, tb

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sb = min(B, 128)
        ts = cpu_steps(sb, 3, 1, threads)
        cpu = {"value": sb * len(ts) / sum(ts), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{len(ts)} steps x {sb} samples of the same step (oracle, fp32 torch CPU ops)"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands (10-bit mantissa, TF32-equivalent) / f32 accumulate",
            "data": "synthetic", "config": workload_config(args, B),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step),
            "cuda_graph": bool(use_graph), "data_parallel_mode": dp_mode, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
            "step_tflops": value * FLOP_PER_SAMPLE / 1e12 / world,
            "step_frac_of_tensor_peak": value * FLOP_PER_SAMPLE / 1e12 / world / pk["tf_sust"],
            "final_loss": final_loss, "top_kernels": table[:6], "batch_sweep": sweep,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MMDYN_BENCH_BATCH", 1024)),
                    help="per-GPU batch (weak scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nccl-in-graph", action="store_true",
                    help="experimental: capture NCCL inside the step graph (hangs on torch 2.11 / NCCL 2.28.9)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-GPU batch sweep (64-4096, 30 graph replays each)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table here (JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
