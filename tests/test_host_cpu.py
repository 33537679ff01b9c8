"""CPU-side checks: the C-ABI library loads and exports every symbol include/mmdyn_b200.h declares
(no compute calls without a GPU), the product path refuses to run without CUDA, and the
data-parallel host logic works across 2 gloo ranks."""
import os
import re
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mmdyn_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "mmdyn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mmdyn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert l.mmdyn_version() >= 100
    assert l.mmdyn_launch_count() >= 0


def test_argument_validation_without_gpu():
    import ctypes as C
    from mmdyn_b200 import lib
    l = lib.load()
    d = lib.IgemmDesc()
    assert l.mmdyn_igemm(C.byref(d), None) < 0
    assert b"null pointer" in l.mmdyn_last_error()
    assert l.mmdyn_adam_flat(None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None) < 0


def test_no_cpu_fallback():
    from mmdyn_b200 import engine
    from mmdyn_b200.pytorch.models.models import setup_model
    m = setup_model("cnn-vae", condition_dim=0, input_dim=4096, architecture="cnn", conditional=False,
                    categorical_conditions=False, latent_size=256)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(2, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        engine.get_arena(m)
    # CVAE shock conditioning is part of the path: same parameter shapes as the reference
    cm = setup_model("cnn-vae", condition_dim=3, input_dim=4096, architecture="cnn", conditional=True,
                     categorical_conditions=False, latent_size=256)
    sd = cm.state_dict()
    assert tuple(sd["encoder.linear_means.weight"].shape) == (256, 515)
    assert tuple(sd["decoder.upsample.0.weight"].shape) == (6400, 259)
    with pytest.raises(NotImplementedError):  # class-label (one-hot) conditions are not
        setup_model("cnn-vae", condition_dim=10, input_dim=4096, architecture="cnn", conditional=True,
                    categorical_conditions=True, latent_size=256)


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodal-dynamics_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_cli_flags_match_reference():
    from mmdyn_b200.pytorch.main import build_parser
    a = build_parser().parse_args([])
    assert (a.problem_type, a.model_name, a.input_type, a.use_pose) == ("seq_modeling", "cnn-mvae", "visual", False)
    assert (a.lr, a.batchsize, a.optimizer, a.num_epochs, a.pose_multiplier) == (0.001, 128, "Adam", 100, 1000)
    assert (a.kl_weight, a.latent_size, a.annealing_epochs, a.conditional, a.mask_loss) == (1.0, 256, 50, False, False)
    b = build_parser().parse_args("--problem-type dyn_modeling --input-type visuotactile --model-name cnn-mvae "
                                  "--use-pose --no-cuda --vis-pose --criterion crossentropy --save-name x".split())
    assert b.problem_type == "dyn_modeling" and b.use_pose and b.no_cuda


def test_bucket_ranges_and_sharding():
    from mmdyn_b200 import engine, parallel
    from mmdyn_b200.pytorch.models.models import setup_model
    m = setup_model("cnn-mvae", cross_modal=True, condition_dim=0, input_dim=4096, architecture="cnn",
                    conditional=False, categorical_conditions=False, latent_size=256, use_pose=True)
    ar = engine.ParamArena(m)
    r = parallel.bucket_ranges(ar.names, ar.offset, ar.numel, ar.total)
    assert list(r) == ["visual_encoder", "visual_decoder", "tactile_encoder", "tactile_decoder", "pose_encoder",
                       "pose_decoder"]
    spans = list(r.values())
    assert spans[0][0] == 0 and spans[-1][1] == ar.total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert sum(p.numel() for p in m.parameters()) == 14058119 <= ar.total
    # sharding by whole sequences
    got = [parallel.shard_rows(7 * 50, 2, k, 50) for k in range(2)]
    assert got == [(0, 200), (200, 350)]
    assert parallel.shard_rows(128, 8, 3) == (48, 64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmdyn_b200 import parallel
    torch.manual_seed(100 + rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    ranges = {"a_decoder": (400, 1000), "a_encoder": (0, 400)}
    sync = parallel.GradSync(flat, ranges, overlap=False)
    sync.begin()
    sync.ready(["a_decoder"])          # decoder gradients are final first
    sync.ready(["a_decoder"])          # idempotent
    sync.finish()                      # reduces the rest, waits
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = torch.allclose(flat, sum(gathered), atol=1e-6)
    # averaged update == oracle-per-shard gradients averaged
    avg = flat / world
    ok = ok and torch.allclose(avg, torch.stack(gathered).mean(0), atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradsync_two_gloo_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_synthetic_dataset_with_shock_field_and_parse_input():
    """`--dataset-path synthetic:N:L:3` (the stand-in for exp 3): 5th data field = one shock force per
    sequence, `SeqModeling.parse_input` picks the first frame's (problems.py:634-673), `_set_condition_dim`
    finds its width the way the reference does (problems.py:676-681)."""
    from mmdyn_b200.pytorch.problems import problems
    from mmdyn_b200.pytorch.utils.datasets import dataset_setup
    d = dataset_setup("synthetic:6:4:3", "seq_modeling", batchsize=2, shuffle=False)
    assert d["seq_length"] == 4 and len(d["train_dataset"].data[0][0][4]) == 3
    data, target = next(iter(d["train_loader"]))
    assert len(data) == 5 and tuple(data[4].shape) == (8, 3) and tuple(data[0].shape) == (8, 3, 64, 64)
    assert torch.equal(data[4][0], data[4][3]) and not torch.equal(data[4][0], data[4][4])  # per sequence
    pr = object.__new__(problems.SeqModeling)
    pr.parameters = {"input_type": "visuotactile"}
    pr._seq_length, pr._device = 4, torch.device("cpu")
    pr.train_dataset = d["train_dataset"]
    xi, ti = pr.parse_input(data, target)
    assert tuple(xi["shock"].shape) == (2, 3) and torch.equal(xi["shock"], data[4][::4])
    pr._set_condition_dim()
    assert pr._condition_dim == 3 and pr._categorical_conditions is False
    # without the field: no shock, condition_dim 0
    d0 = dataset_setup("synthetic:6:4", "seq_modeling", batchsize=2, shuffle=False)
    data0, target0 = next(iter(d0["train_loader"]))
    xi0, _ = pr.parse_input(data0, target0)
    assert len(data0) == 4 and xi0["shock"] is None
    pr.train_dataset = d0["train_dataset"]
    pr._set_condition_dim()
    assert pr._condition_dim == 0


def test_bench_clock_sampler_survives_without_a_gpu():
    """bench.py's clock sampler must never take the bench down: without NVML / nvidia-smi it reports
    'unavailable' instead of raising."""
    import importlib.util
    import time
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.3)
    s.stop_flag = True
    s.join(timeout=10)
    out = s.summary()
    assert "sm_mhz" in out and "reasons" in out
    cfg = bench.workload_config(type("A", (), {"gpus": 2})(), 1024)
    assert cfg["global_batch"] == 2048 and cfg["parallelism"] == "dp2" and "workload" in cfg


def test_plan_phase_merging_knob(monkeypatch):
    """MMDYN_MERGE_PHASES=0 keeps the 4 sub-pixel phases of the stride-2 layers as separate 4-tap GEMMs;
    both forms index the same weights (every arena element appears in both packings)."""
    import importlib
    from mmdyn_b200 import plan
    merged = plan.deconv_s2_plan("deconv2", 1000, 128, 64, 8)
    monkeypatch.setenv("MMDYN_MERGE_PHASES", "0")
    plan0 = importlib.reload(plan)
    try:
        split = plan0.deconv_s2_plan("deconv2", 1000, 128, 64, 8)
        assert merged.fwd.n_phases == 1 and merged.fwd.ntaps == 9 and merged.fwd.N == 4 * 64
        assert split.fwd.n_phases == 4 and split.fwd.ntaps == 4 and split.fwd.N == 64
        a = set(merged.idx_fwd[merged.idx_fwd >= 0].tolist())
        b = set(split.idx_fwd[split.idx_fwd >= 0].tolist())
        assert a == b and len(a) == 128 * 64 * 16
    finally:
        monkeypatch.delenv("MMDYN_MERGE_PHASES")
        importlib.reload(plan)


def test_dyn_modeling_is_parsed_globally_then_sharded_by_whole_sequences():
    """DynModeling.parse_input rolls the targets over dim 0 (problems.py:785-798), so a
    shard's last row needs the next shard's first row (and the very last pose target wraps to row 0).
    parallel.shard_batch slices the globally parsed batch; the shards concatenate back to the
    single-process batch, whereas parsing each shard locally changes exactly one pose-target row per rank."""
    from mmdyn_b200 import parallel
    from mmdyn_b200.pytorch.problems import problems
    S, L, world = 4, 5, 2
    g = torch.Generator().manual_seed(3)
    n = S * L
    data = [torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 7, generator=g),
            torch.ones(n, 2)]
    target = [torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 3, 4, 4, generator=g), torch.rand(n, 7, generator=g),
              (torch.rand(n, 3, 4, 4, generator=g) > 0.5).float()]
    pr = object.__new__(problems.DynModeling)
    pr.parameters = {"input_type": "visuotactile"}
    pr._seq_length, pr._device = L, torch.device("cpu")
    xi, ti = pr.parse_input([t.clone() for t in data], [t.clone() for t in target])
    parts = [parallel.shard_batch((xi, ti), world, r, L) for r in range(world)]
    for r, (xs, ts) in enumerate(parts):
        a, b = parallel.shard_rows(n, world, r, L)
        assert (a, b) == (r * 10, r * 10 + 10) and xs["model_input"][0].shape[0] == 10
        assert torch.equal(xs["model_input"][1], xi["model_input"][1][a:b])
        assert torch.equal(ts["target_object_pose"][0], ti["target_object_pose"][0][a:b])
    cat = torch.cat([p[1]["target_object_pose"][0] for p in parts])
    assert torch.equal(cat, ti["target_object_pose"][0])
    # the alternative (each rank parses its own rows) differs in the last pose-target row of every shard
    for r in range(world):
        a, b = parallel.shard_rows(n, world, r, L)
        _, tl = pr.parse_input([t[a:b].clone() for t in data], [t[a:b].clone() for t in target])
        diff = (tl["target_object_pose"][0] != ti["target_object_pose"][0][a:b]).any(dim=1)
        assert diff.nonzero().flatten().tolist() == [b - a - 1]
        assert torch.equal(tl["target_output"][0], ti["target_output"][0][a:b])  # image targets stay local


def test_workspace_never_recycles_a_buffer():
    """Captured CUDA graphs keep raw pointers into Workspace buffers (advisor finding, round 1): a request for
    a known name with another shape must allocate next to the old buffer, not replace it."""
    from mmdyn_b200 import engine
    ws = engine.Workspace(torch.device("cpu"))
    a = ws("act", (8, 4), torch.float32)
    s1 = ws("sums", (2, 3), torch.float32, zero="step")
    b = ws("act", (4, 4), torch.float32)          # another batch size, same name
    s2 = ws("sums", (1, 3), torch.float32, zero="step")
    a2 = ws("act", (8, 4), torch.float32)
    assert a2.data_ptr() == a.data_ptr() and b.data_ptr() != a.data_ptr()
    assert ws("sums", (2, 3), torch.float32, zero="step").data_ptr() == s1.data_ptr() != s2.data_ptr()
    assert ws.bufs["act"] is a2
    s1.fill_(3.0)
    s2.fill_(4.0)
    ws.begin_step()                                # one memset clears every zero="step" buffer
    assert float(s1.abs().sum()) == 0.0 and float(s2.abs().sum()) == 0.0
    z = ws("z", (5,), torch.float32, zero=True)
    z.fill_(1.0)
    assert float(ws("z", (5,), torch.float32, zero=True).sum()) == 0.0
    n = ws.nbytes()
    ws.trim()
    assert n > 0 and ws.nbytes() == 0 and not ws.bufs


def test_device_noise_state_round_trip():
    from mmdyn_b200 import noise
    src = noise.DeviceNoise(seed=123)
    assert src.state_dict() == {"seed": 123, "counter": 0}
    src._counter("cpu").fill_(987654321012)
    st = src.state_dict()
    back = noise.DeviceNoise.from_state_dict(st, "cpu")
    assert back.seed == 123 and int(back.ctr.item()) == 987654321012


def test_bench_dyn_shards_equal_the_global_parse():
    """bench.py --problem dyn builds every rank's rows locally (its sequences + the next rank's first one); the
    result must be the rows DynModeling.parse_input gives on the WHOLE step batch, incl. the wrap-around row."""
    import bench
    from oracle import mmdyn_oracle as orc
    world, S, L = 3, 2, 4
    shards = [bench.dyn_shard(r, world, S, L, seed=5) for r in range(world)]
    # global batch from the same per-sequence generator
    def seq(i):
        g = torch.Generator().manual_seed(5 * 1000003 + i)
        r = lambda *s_: torch.rand(*s_, generator=g)
        return r(L, 3, 64, 64), r(L, 3, 64, 64), r(L, 7), r(1, 3, 64, 64), r(1, 3, 64, 64)
    parts = [seq(i) for i in range(world * S)]
    data = [torch.cat([p[k] for p in parts]) for k in range(3)] + [torch.ones(world * S * L, 2)]
    target = [torch.cat([p[3].expand(L, -1, -1, -1) for p in parts]), torch.cat([p[4].expand(L, -1, -1, -1) for p in parts]),
              data[2].clone(), torch.ones(world * S * L, 3, 64, 64)]
    inp, tgt = orc.dyn_parse_input(data, target, L, "visuotactile")
    n = S * L
    for r, (x, t) in enumerate(shards):
        sl = slice(r * n, (r + 1) * n)
        assert torch.equal(x[0], inp["model_input"][0][sl]) and torch.equal(x[2], inp["input_object_pose"][0][sl])
        assert torch.equal(t[0], tgt["target_output"][0][sl]) and torch.equal(t[1], tgt["target_output"][1][sl])
        assert torch.equal(t[2], tgt["target_object_pose"][0][sl]), r  # bare roll; last rank wraps to row 0


def test_packer_stages_follow_the_order_a_step_needs_them():
    """engine.ModelPacker lays the fp16 weight copies out in three contiguous stages — forward copies of the encoder-side
    networks, forward copies of the decoders, data-gradient copies — so that a staged refresh can pack the later stages on
    a side stream.  Views must tile the arena without overlap, 16-byte aligned."""
    from types import SimpleNamespace
    from mmdyn_b200 import engine, plan

    def layer(nf, nd, name):
        mk = lambda n: torch.arange(n, dtype=torch.int32).view(8, -1) if n else None
        return SimpleNamespace(idx_fwd=mk(nf), idx_dgrad=mk(nd), bias_idx=None, lp=SimpleNamespace(name=name),
                               Wf=None, Wd=None, bias=None)

    enc = SimpleNamespace(layers=[layer(64, 128, "e0"), layer(256, 0, "e1")])
    dec = object.__new__(engine.DecoderExec)
    dec.layers = [layer(512, 1024, "d0")]
    pose = SimpleNamespace(layers=[layer(192, 192, "p0")])
    pk = engine.ModelPacker([enc, dec, pose], torch.device("cpu"))
    s0, s1 = 64 + 256 + 192, 64 + 256 + 192 + 512
    assert pk.stage_off == [0, s0, s1, s1 + 128 + 1024 + 192]
    assert pk.idx.numel() == pk.W.numel() == pk.stage_off[-1]
    spans = sorted((o, o + sh[0] * sh[1], attr, pl.lp.name) for pl, attr, o, sh in pk.slots)
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == pk.stage_off[-1]
    assert all(o % 8 == 0 for o, *_ in spans)
    for o, e, attr, name in spans:
        stage = 2 if attr == "Wd" else (1 if name == "d0" else 0)
        assert pk.stage_off[stage] <= o and e <= pk.stage_off[stage + 1], (name, attr)
    assert dec.layers[0].Wf.shape == (8, 64) and dec.layers[0].Wf.data_ptr() == pk.W[s0:].data_ptr()
    # images per GEMM tile (fused BatchNorm statistics need whole tiles per group): same rule as the launcher
    assert plan.tile_images(plan.conv_s2_plan("c2", 0, 32, 64, 32).fwd) == 1      # 16 x 8 pixels of one image
    assert plan.tile_images(plan.conv_s2_plan("c3", 0, 64, 128, 16).fwd) == 2     # 8 x 8 pixels x 2 images
    assert plan.tile_images(plan.conv_k4s1p0_plan("c4", 0, 128, 256, 8).fwd) == 128  # one pixel x 128 images
    assert plan.tile_images(plan.deconv_k4s1p0_plan("d1", 0, 256, 128, 5).fwd) == 128
