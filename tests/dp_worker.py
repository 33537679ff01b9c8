"""Data-parallel parity worker — launched by tests/test_dp_gpu.py (and by hand) as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dp_worker.py

One process per GPU, NCCL.  Parity definition: N ranks == the ORACLE run independently on each
rank's shard with identical weights, gradients averaged, one optimizer step.  Checked for
  1. seq_modeling cnn-mvae (3 passes) and cnn-mvae + pose (7 passes): eager step, bucketed all-reduce
     launched from the backward (parallel.GradSync);
  2. dyn_modeling (BASELINE.json configs[3]): the global batch is parsed once (roll / fix-up / wrap-around row
     of problems.py:765-803) and sharded by whole sequences; --mask-loss variant without the pose expert and
     the pose variant;
  3. the CUDA-graph step in its data-parallel forms (split graphs + flat all-reduce; bucketed all-reduce
     captured inside the graph when MMDYN_TEST_NCCL_IN_GRAPH=1; the peer-memory exchange kernel when
     available): same gradients as the eager path, bit-identical parameters on every rank afterwards.
Exits non-zero on the first failed assertion (the launcher then fails the pytest).
"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import mmdyn_oracle as orc  # noqa: E402  (test infrastructure: the checker)

KW = dict(condition_dim=0, input_dim=4096, architecture="cnn", conditional=False, categorical_conditions=False,
          latent_size=256)


def nrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def log(*a):
    if dist.get_rank() == 0:
        print(*a, flush=True)


def make(use_pose, seed, dev):
    from mmdyn_b200.pytorch.models.models import setup_model
    torch.manual_seed(seed)  # identical replicas on every rank
    m = setup_model("cnn-mvae", cross_modal=True, use_pose=use_pose, **KW)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m.to(dev), sd


def global_batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    return dict(v=r(n, 3, 64, 64), t=r(n, 3, 64, 64), p=r(n, 7), tv=r(n, 3, 64, 64), tt=r(n, 3, 64, 64), tp=r(n, 7))


def oracle_shard_grads(sd, pkeys, x, t, klw, use_pose, seed, dev, loss_mask=None):
    """Oracle on THIS rank's shard; gradients averaged over ranks (the all-reduce is only the test's plumbing)."""
    passes = orc.MVAE_PASSES_POSE if use_pose else orc.MVAE_PASSES_NOPOSE
    sd_o = copy.deepcopy(sd)
    for k in pkeys:
        sd_o[k].requires_grad_(True)
    g = torch.Generator().manual_seed(seed)
    noises = [orc.draw_pass_noise(x[0].shape[0], hv, ht, generator=g) for (hv, ht, hp) in passes]
    _, loss, _ = orc.evaluate_mvae(sd_o, x, t, klw, 1000.0, use_pose, noises, loss_mask=loss_mask)
    loss.backward()
    flat = torch.cat([sd_o[k].grad.reshape(-1) for k in pkeys]).to(dev)
    dist.all_reduce(flat)
    return flat / dist.get_world_size(), float(loss.detach())


def arena_params_grad(model):
    return torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])


def check_replicas_identical(model, what):
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    mine = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat[::997].double().pow(2).sum()])
    allv = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allv, mine)
    for o in allv[1:]:
        assert torch.equal(o, allv[0]), f"{what}: replicas diverged: {[v.tolist() for v in allv]}"


def seq_case(use_pose, dev, rank, world):
    from mmdyn_b200 import engine, noise, optim, parallel
    klw, Bl = 0.02, 4
    model, sd = make(use_pose, 9, dev)
    pkeys = [k for k, _ in model.named_parameters()]
    d = global_batch(Bl * world, 2)
    a, b = parallel.shard_rows(Bl * world, world, rank)
    x = [d["v"][a:b], d["t"][a:b]] + ([d["p"][a:b]] if use_pose else [])
    t = [d["tv"][a:b], d["tt"][a:b]] + ([d["tp"][a:b]] if use_pose else [])
    g_o, _ = oracle_shard_grads(sd, pkeys, x, t, klw, use_pose, 30 + rank, dev)
    eng = engine.StepEngine(model, "mvae", use_pose=use_pose, noise_src=noise.HostNoise(torch.Generator().manual_seed(30 + rank)))
    opt = optim.FusedAdam(model, lr=1e-3)
    arena = engine.get_arena(model, dev)
    sync = parallel.attach(eng, opt, arena, overlap=True)
    assert opt.grad_prescale == 1.0 / world
    opt.zero_grad()
    sync.begin()
    _, loss = eng.evaluate([v.to(dev) for v in x], [v.to(dev) for v in t], klw, want_outputs=False)
    loss.backward()            # bucket hooks launch the all-reduces on the side stream
    sync.finish()
    torch.cuda.synchronize()
    g_d = arena_params_grad(model) / world
    e = nrel(g_d, g_o)
    log(f"[dp] seq_modeling pose={use_pose}: {world}-rank averaged gradients vs oracle-per-shard average: rel {e:.3e}")
    assert e < (2e-2 if use_pose else 5e-3), e
    opt.step()
    opt.check_finite(float(loss))
    check_replicas_identical(model, f"seq pose={use_pose}")


def dyn_case(use_pose, dev, rank, world):
    """configs[3]: dyn_modeling, all S*L frames of the step, targets by roll/fix-up, sharded by whole sequences."""
    from mmdyn_b200 import engine, noise, optim, parallel
    from mmdyn_b200.pytorch.problems import problems
    klw, L = 0.02, 3
    S = 2 * world
    n = S * L
    model, sd = make(use_pose, 5, dev)
    pkeys = [k for k, _ in model.named_parameters()]
    d = global_batch(n, 8)
    gm = torch.Generator().manual_seed(4)
    seg = (torch.rand(n, 3, 64, 64, generator=gm) > 0.4).float()
    avail = torch.ones(n, 2)
    data = [d["v"], d["t"], d["p"], avail]
    target = [d["tv"], d["tt"], d["tp"], seg]
    masked = not use_pose  # --mask-loss + --use-pose does not broadcast in the reference either (problems.py:446)
    # ---- oracle: parse globally, shard, evaluate the shard ----
    inp_o, tgt_o = orc.dyn_parse_input(data, target, L, "visuotactile")
    a, b = parallel.shard_rows(n, world, rank, L)
    assert (b - a) % L == 0
    x_o = [t_[a:b] for t_ in inp_o["model_input"]] + ([inp_o["input_object_pose"][0][a:b]] if use_pose else [])
    t_o = [t_[a:b] for t_ in tgt_o["target_output"]] + ([tgt_o["target_object_pose"][0][a:b]] if use_pose else [])
    m_o = tgt_o["loss_mask"][a:b] if masked else None
    g_o, loss_o = oracle_shard_grads(sd, pkeys, x_o, t_o, klw, use_pose, 50 + rank, dev, loss_mask=m_o)
    # ---- product: DynModeling.parse_input on the whole step batch, then shard_batch ----
    pr = object.__new__(problems.DynModeling)
    pr.parameters = {"model_name": "cnn-mvae", "input_type": "visuotactile", "use_pose": use_pose, "mask_loss": masked}
    pr._kl_weight, pr._pose_multiplier, pr._conditional = klw, 1000.0, False
    pr._model, pr._cross_modal, pr._engine = model, True, None
    pr._seq_length, pr._device = L, dev
    inputs, targets = pr.parse_input(data, target)
    inputs, targets = parallel.shard_batch(inputs, world, rank, L), parallel.shard_batch(targets, world, rank, L)
    assert torch.equal(inputs["model_input"][0].cpu(), x_o[0]) and torch.equal(targets["target_output"][1].cpu(), t_o[1])
    assert torch.equal(targets["target_object_pose"][0].cpu(), tgt_o["target_object_pose"][0][a:b])  # incl. the wrap-around row
    model.noise = noise.HostNoise(torch.Generator().manual_seed(50 + rank))
    eng = pr._get_engine()
    opt = optim.FusedAdam(model, lr=1e-3)
    arena = engine.get_arena(model, dev)
    sync = parallel.attach(eng, opt, arena, overlap=True)
    opt.zero_grad()
    sync.begin()
    _, loss = pr._evaluate_model(inputs, targets)
    loss.backward()
    sync.finish()
    torch.cuda.synchronize()
    el = abs(float(loss) - loss_o) / abs(loss_o)
    e = nrel(arena_params_grad(model) / world, g_o)
    log(f"[dp] dyn_modeling pose={use_pose} masked={masked}: shard loss rel {el:.2e}, averaged gradients rel {e:.3e}")
    assert el < 1e-4 and e < (2e-2 if use_pose else 5e-3), (el, e)
    opt.step()
    check_replicas_identical(model, f"dyn pose={use_pose}")


def graph_case(dev, rank, world):
    """The CUDA-graph step, data parallel: every available exchange scheme gives the eager path's summed
    gradients and leaves bit-identical replicas.  (No pose expert here: its ReLU units flip under the
    replay-to-replay fp32-atomics noise and would blur the comparison, see tests/test_parity2_gpu.py.)"""
    from mmdyn_b200 import engine, noise, optim, parallel
    klw, Bl = 0.02, 8
    model, sd = make(False, 13, dev)
    d = global_batch(Bl * world, 21)
    a, b = parallel.shard_rows(Bl * world, world, rank)
    x = [d["v"][a:b].to(dev), d["t"][a:b].to(dev)]
    t = [d["tv"][a:b].to(dev), d["tt"][a:b].to(dev)]
    src = noise.DeviceNoise(seed=900 + rank)
    eng = engine.StepEngine(model, "mvae", use_pose=False, noise_src=src)
    opt = optim.FusedAdam(model, lr=1e-3)
    arena = engine.get_arena(model, dev)
    opt.grad_prescale = 1.0 / world
    w0 = arena.flat.clone()

    def reset():  # same weights, fresh optimizer moments, same noise stream for every scheme
        arena.flat.copy_(w0)
        arena.bump()
        for buf in opt._bufs or ():
            buf.zero_()
        if opt._bufs is not None:
            opt._step_dev.zero_()
        src._counter(dev).zero_()

    # reference: eager step, flat all-reduce, one Adam step
    opt._arena()
    reset()
    opt.zero_grad()
    _, l0 = eng.evaluate(x, t, klw, need_grad=True, autograd=False, want_outputs=False)
    eng.backward()
    dist.all_reduce(arena.grad)
    torch.cuda.synchronize()
    g_ref, l_ref = arena.grad.clone(), float(l0)
    opt.step()
    torch.cuda.synchronize()
    w_ref = arena.flat.clone()
    schemes = ["split"]
    if os.environ.get("MMDYN_TEST_NCCL_IN_GRAPH") == "1":
        schemes.append("nccl_in_graph")
    if hasattr(parallel, "PeerExchange") and os.environ.get("MMDYN_TEST_NO_PEER") is None:
        schemes.append("peer")
    for scheme in schemes:
        reset()
        torch.cuda.synchronize()
        dist.barrier()
        if scheme == "split":
            g = engine.GraphedTrainStep(eng, opt, x, t, klw, split_optimizer=True)
        elif scheme == "nccl_in_graph":
            sync = parallel.attach(eng, opt, arena, overlap=True)
            g = engine.GraphedTrainStep(eng, opt, x, t, klw, split_optimizer=True, grad_sync=sync)
        else:
            px = parallel.PeerExchange(arena, opt)
            g = engine.GraphedTrainStep(eng, opt, x, t, klw, peer_exchange=px)
        reset()
        g.run()
        if scheme == "split":
            dist.all_reduce(arena.grad)
        torch.cuda.synchronize()
        el = abs(float(g.loss) - l_ref) / abs(l_ref)
        if scheme != "peer":
            e = nrel(arena.grad, g_ref)
            log(f"[dp] graph step, scheme {scheme}: loss rel {el:.1e}, summed gradients vs eager rel {e:.2e}")
            # not bit-equal: fp32 atomics reorder from launch to launch and a few fp16 roundings flip behind them
            assert el < 2e-6 and e < 1e-3, (scheme, el, e)
            g.apply()
        else:
            # the fused exchange leaves the local gradients untouched: rebuild what it must have computed from them
            gsum = arena.grad.clone()
            dist.all_reduce(gsum)
            gavg = gsum / world
            want = w0 - 1e-3 * (0.1 * gavg / 0.1) / ((0.001 * gavg * gavg).sqrt() / (0.001 ** 0.5) + 1e-8)  # Adam, step 1
            # entries with |g| ~ eps = 1e-8 amplify the summation-order difference between NCCL's all-reduce and the
            # kernel's rank-order sum (step 1 of Adam is lr * g / (|g| + eps)): judge the well-conditioned entries
            sel = gavg.abs() > 1e-5
            e = (arena.flat - want)[sel].abs().max().item()
            e_all = (arena.flat - want).abs().max().item()
            log(f"[dp] graph step, scheme peer: parameters vs Adam(all-reduced local gradients): max abs diff {e:.2e} "
                f"over the {int(sel.sum())} entries with |g| > 1e-5 ({e_all:.2e} over all, lr = 1e-3)")
            assert e < 2e-6 and e_all < 1.01e-3, (e, e_all)
        torch.cuda.synchronize()
        ew = nrel(arena.flat - w0, w_ref - w0)
        log(f"[dp] graph step, scheme {scheme}: Adam displacement vs eager rel {ew:.2e} (loss rel {el:.1e})")
        # step 1 of Adam is lr * sign(g) wherever |g| >> eps: entries whose tiny gradient changes sign under the
        # replay noise move by 2 lr, the rest agree exactly
        assert el < 2e-6 and ew < 5e-2, (scheme, el, ew)
        check_replicas_identical(model, f"graph {scheme}")
        eng.bucket_hook = None
        del g


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    sys.stdout.flush()
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    which = os.environ.get("MMDYN_DP_CASES", "seq,dyn,graph").split(",")
    if "seq" in which:
        seq_case(False, dev, rank, world)
        seq_case(True, dev, rank, world)
    if "dyn" in which:
        dyn_case(False, dev, rank, world)
        dyn_case(True, dev, rank, world)
    if "graph" in which:
        graph_case(dev, rank, world)
    dist.barrier()
    log(f"[dp] all data-parallel parity cases passed on {world} ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
