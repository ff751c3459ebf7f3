"""CPU tests of the host-side logic: ray sharding over 2 gloo ranks, the interleaved weight packing, the model mirror's
state-dict contract (against the real reference classes when /root/reference is present), the CPU frame port."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_worker(rank, world, port, H, W, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radnerf_b200.sharding import FrameSharder
    sh = FrameSharder(H, W, world, rank, torch.device("cpu"))
    full = torch.arange(H * W * 3, dtype=torch.float32).view(H * W, 3)
    local = sh.shard(full)
    assert local.shape == (H * W // world, 3)
    # each rank "renders" its rows: image = f(pixel id); the gathered image must be f(all pixels) in pixel order
    img = sh.gather(local * 2 + 1)
    ok = torch.equal(img, full * 2 + 1)
    rows = (sh.ids // W).unique()
    out[rank] = (ok, rows.tolist())
    dist.destroy_process_group()


def test_frame_sharding_two_ranks_gloo():
    H, W, world = 32, 16, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 1000
    mp.spawn(_shard_worker, args=(world, port, H, W, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    r0, r1 = set(out[0][1]), set(out[1][1])
    assert r0 | r1 == set(range(H)) and not (r0 & r1)
    # interleaved 8-row tiles: rank 0 owns rows 0-7, 16-23; rank 1 owns 8-15, 24-31
    assert r0 == set(range(0, 8)) | set(range(16, 24))


def _toy_model():
    torch.manual_seed(0)
    m = torch.nn.Module()
    m.table = torch.nn.Parameter(torch.randn(512, 2) * 0.1)   # >= table_numel below: own asynchronous all-reduce
    m.lin = torch.nn.Linear(2, 3)                              # small parameters: flat bucket
    m.unused = torch.nn.Parameter(torch.zeros(4))              # never receives a gradient
    return m


def _toy_loss(m, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, 512, (64,), generator=g)
    target = torch.randn(64, 3, generator=g)
    return ((m.lin(m.table[idx]) - target) ** 2).mean()


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radnerf_b200.train import GradSync
    m = _toy_model()
    sync = GradSync(m.parameters(), table_numel=1024)
    assert [p.numel() for p in sync.big] == [1024] and len(sync.small) == 3
    got = []
    for step in range(2):   # two steps: the hooks must keep firing and grads must not accumulate across steps
        for p in m.parameters():
            if p.grad is not None:
                p.grad.zero_()
        sync.begin_step()
        _toy_loss(m, 100 * step + rank).backward()   # every rank draws its own batch
        sync.finish()
        got.append([None if p.grad is None else p.grad.clone() for p in m.parameters()])
    out[rank] = (got, sync.bytes_last)
    dist.destroy_process_group()


def test_gradsync_two_ranks_gloo_equals_mean_of_local_grads():
    """Data-parallel gradient exchange: after GradSync.finish() every rank must hold the MEAN over ranks of the per-rank
    gradients (table via its own hook-driven all-reduce, small parameters via the flat bucket); unused parameters stay None."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() + 17) % 1000
    mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
    for step in range(2):
        want = None
        for r in range(world):
            m = _toy_model()
            _toy_loss(m, 100 * step + r).backward()
            g = [None if p.grad is None else p.grad / world for p in m.parameters()]
            want = g if want is None else [a if b is None else a + b for a, b in zip(want, g)]
        for r in range(world):
            for a, b in zip(out[r][0][step], want):
                assert (a is None) == (b is None)
                if a is not None:
                    assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    assert out[0][1] == (1024 + 6 + 3) * 4   # bytes exchanged per step: the table + the flat bucket (lin.weight, lin.bias)


class _OccupancyStub:
    """stands in for NeRFNetwork: an update that consumes the RNG streams update_extra_state consumes and derives its grid from
    the draws and the step counters"""

    def __init__(self, rank):
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32)
        self.step_counter[:4, 0] = torch.tensor([100, 120, 90, 110]) * (rank + 1)     # ranks emitted different sample counts
        self.step_counter[:4, 1] = 64
        self.grid = None

    def update_extra_state(self):
        import random
        k = random.randint(0, 599)
        self.grid = torch.rand(32) + k
        self.mean_count = int(self.step_counter[:4, 0].sum().item() / 4)
        return k


def _occupancy_worker(rank, world, port, out):
    import random
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radnerf_b200.train import update_extra_state_replicated
    torch.manual_seed(1000 + rank)       # every rank has its own RNG streams (its own rays) ...
    random.seed(2000 + rank)
    m = _OccupancyStub(rank)
    before = (torch.rand(3), random.random())
    torch.manual_seed(1000 + rank)
    random.seed(2000 + rank)
    k = update_extra_state_replicated(m)
    after = (torch.rand(3), random.random())      # ... which the replicated update must leave where they were
    out[rank] = (k, m.grid, m.mean_count, m.step_counter.clone(), torch.equal(before[0], after[0]) and before[1] == after[1])
    dist.destroy_process_group()


def test_replicated_occupancy_update_two_ranks_gloo():
    """SURVEY 8(e) "Occupancy update" / "mean_count": same draws and the same (max-reduced) counters on every rank, per-rank
    RNG streams untouched"""
    world = 2
    out = mp.Manager().dict()
    port = 29500 + (os.getpid() + 131) % 1000
    mp.spawn(_occupancy_worker, args=(world, port, out), nprocs=world, join=True)
    (k0, g0, mc0, sc0, kept0), (k1, g1, mc1, sc1, kept1) = out[0], out[1]
    assert k0 == k1 and torch.equal(g0, g1)                     # identical random draws -> identical grids
    assert mc0 == mc1 == (200 + 240 + 180 + 220) // 4           # counters max-reduced: the busier rank's counts
    assert torch.equal(sc0, sc1) and kept0 and kept1


def test_sharder_rejects_unbalanced_split():
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    from radnerf_b200.sharding import FrameSharder
    with pytest.raises(ValueError):
        FrameSharder(20, 16, 4, 0, torch.device("cpu"))
    assert FrameSharder(20, 16, 1, 0, torch.device("cpu")).shard(torch.zeros(320, 3)).shape == (320, 3)


def test_interleaved_weight_packing_matches_the_umma_layout():
    """il_pack must place W[r, k] at byte (r/8)*(128*K/8) + (k/8)*128 + (r%8)*16 + (k%8)*2 (csrc/umma.cuh il_offset)."""
    from radnerf_b200.frame import il_pack
    for n, k, n_pad, k_pad in ((64, 32, 64, 32), (2, 64, 16, 64), (65, 64, 80, 64), (64, 42, 64, 48), (32, 74, 32, 80)):
        W = torch.randn(n, k)
        blob = il_pack(W, n_pad, k_pad).view(torch.int16).numpy()
        ref = W.half().view(torch.int16).numpy()
        for r, c in ((0, 0), (min(1, n - 1), 0), (min(7, n - 1), 7), (min(8, n - 1), 0), (n - 1, k - 1), (n // 2, k // 3)):
            off = (r // 8) * (128 * k_pad // 8) + (c // 8) * 128 + (r % 8) * 16 + (c % 8) * 2
            assert blob[off // 2] == ref[r, c]
        assert blob.size == n_pad * k_pad
        if n_pad > n:  # padding rows are zero
            off = (n // 8) * (128 * k_pad // 8) + (n % 8) * 16
            assert blob[off // 2] == 0


def _stub_reference_imports():
    for name in ("trimesh", "tensorboardX", "matplotlib", "matplotlib.pyplot", "mcubes", "imageio", "lpips", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "torch_ema" not in sys.modules:
        m = types.ModuleType("torch_ema")
        m.ExponentialMovingAverage = object
        sys.modules["torch_ema"] = m


@pytest.mark.skipif(not os.path.isdir("/root/reference/nerf"), reason="reference tree not present")
def test_model_mirror_has_the_reference_state_dict():
    """Construct the REAL reference NeRFNetwork on top of our drop-in operator packages (CPU, no kernels run) and compare
    its state-dict keys/shapes and encoder geometry with radnerf_b200.model.NeRFNetwork."""
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    sys.path.append("/root/reference")
    _stub_reference_imports()
    from nerf.network import NeRFNetwork as RefNet  # the reference's class, importing OUR gridencoder/raymarching/...
    import raymarching
    assert raymarching.__file__.startswith(os.path.join(ROOT, "rad-nerf_b200")), "drop-in packages must shadow the reference's"
    from radnerf_b200.model import NeRFNetwork, Options
    opt = Options(torso=True, smooth_lips=True)
    ns = types.SimpleNamespace(**{**vars(opt), "test_train": False})
    ref = RefNet(ns)
    ours = NeRFNetwork(opt)
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert set(sd_r) == set(sd_o), (set(sd_r) ^ set(sd_o))
    for k in sd_r:
        assert sd_r[k].shape == sd_o[k].shape and sd_r[k].dtype == sd_o[k].dtype, k
    ours.load_state_dict(sd_r)  # a reference checkpoint loads as is
    assert ref.encoder.offsets.tolist() == ours.encoder.offsets.tolist()
    assert repr(ref.encoder) == repr(ours.encoder) and repr(ref.torso_deform_encoder) == repr(ours.torso_deform_encoder)
    assert ref.in_dim == 32 and ref.in_dim_dir == 16 and ref.torso_deform_in_dim == 42 and ref.pose_in_dim == 54


def test_cpu_port_renders_a_frame():
    """The oracle-backed CPU port (bench.py's cpu_baseline / --impl reference arm) renders a small head+torso frame."""
    sys.path.insert(0, ROOT)
    import bench
    fps, threads, nsamp, timed = bench.cpu_frame_rate(48, 1)
    assert fps > 0 and threads >= 1 and nsamp > 0


def test_posemath_roundtrip():
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    from radnerf_b200.posemath import matrix_to_euler_xyz, euler_xyz_to_matrix, convert_poses
    e = torch.tensor([[0.3, -0.4, 1.1], [-1.2, 0.2, 0.5]])
    assert torch.allclose(matrix_to_euler_xyz(euler_xyz_to_matrix(e)), e, atol=1e-6)
    P = torch.eye(4)[None].repeat(2, 1, 1)
    P[:, :3, :3] = euler_xyz_to_matrix(e)
    P[:, :3, 3] = torch.tensor([[0.07, 3.38, -0.23], [1, 2, 3]])
    out = convert_poses(P)
    assert out.shape == (2, 6) and torch.allclose(out[:, :3], e, atol=1e-6) and torch.allclose(out[:, 3:], P[:, :3, 3])


def test_sequence_slices_partition_the_frames():
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    from radnerf_b200.stream import sequence_slice
    for n in (0, 1, 7, 200, 201):
        for world in (1, 2, 3, 8):
            spans = [sequence_slice(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))             # contiguous, in order
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)   # balanced, larger slices first


@pytest.mark.gpu
def test_pack_inputs_layout_matches_the_graph_input_block():
    """[pose 4x4 | pose6 | eye | pad | audio window]: the layout render_frame's static input block and FrameStreamer rely on"""
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    if not torch.cuda.is_available():
        pytest.skip("pinned host memory needs a CUDA driver")
    from radnerf_b200.stream import pack_inputs
    pose = np.arange(16, dtype=np.float32).reshape(4, 4)
    auds = np.arange(8 * 44 * 16, dtype=np.float32).reshape(8, 44, 16) + 100
    p = pack_inputs(pose, auds, pose6=np.arange(6) + 50, eye=np.array([[0.25]], np.float32)).numpy()
    assert p.shape == (24 + 8 * 44 * 16,)
    assert np.array_equal(p[:16], pose.reshape(-1)) and np.array_equal(p[16:22], np.arange(6) + 50) and p[22] == 0.25 and p[23] == 0
    assert np.array_equal(p[24:], auds.reshape(-1))


def test_config1_case(oracle):
    """BASELINE configs[0] (the reference's CPU-runnable case: march -> GridEncoder fwd+bwd -> composite_rays_train fwd+bwd) at a
    small size on the CPU restatement: the pipeline is self-consistent (exact packing of the marcher's output, weights in
    [0,1], gradient sums) and independent of the marcher's slot order (tools/config1.py compares GPU and CPU runs that way)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("config1", os.path.join(ROOT, "tools", "config1.py"))
    c1 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(c1)
    c = c1.inputs(4096)
    res, timing = c1.run_cpu(c, reps=1)
    m, rays = timing["samples"], res["rays"]
    assert 0 < m <= 4096 * c1.MAX_STEPS and timing["cores"] >= 1
    assert int(res["counter"][0]) == m == int(rays[:, 2].sum()) and int(res["counter"][1]) == 4096
    order = np.argsort(rays[:, 1], kind="stable")                       # offsets tile [0, m) exactly
    assert np.array_equal(np.cumsum(rays[order, 2]) - rays[order, 2], rays[order, 1])
    assert res["weights_sum"].min() >= 0 and res["weights_sum"].max() <= 1 + 1e-6
    assert res["feat"].shape == (m, 32) and np.isfinite(res["g_table"]).all()
    # the table gradient distributes exactly the incoming gradient: per level, sum over rows == sum over samples (weights sum to 1)
    s = c1.sample_inputs(c, rays, m)
    offs = c["offsets"]
    for lvl in (0, 7, 15):
        want = s["d_feat"][:, 2 * lvl:2 * lvl + 2].astype(np.float64).sum(0)
        got = res["g_table"][offs[lvl]:offs[lvl + 1]].sum(0)
        assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    # a run whose slots are permuted ray by ray compares equal in canonical order
    perm = np.random.default_rng(3).permutation(4096)
    shuffled = dict(res)
    new_rays, pieces, off = np.empty_like(rays), {k: [] for k in ("xyzs", "dirs", "deltas", "feat", "g_sigmas", "g_rgbs", "g_ambient")}, 0
    for j, i in enumerate(perm):
        rid, o, k = rays[i]
        new_rays[j] = (rid, off, k)
        for key in pieces:
            pieces[key].append(res[key][o:o + k])
        off += k
    shuffled["rays"] = new_rays
    for key in pieces:
        shuffled[key] = np.concatenate(pieces[key])
    # per-ray outputs are addressed by ray id (raymarching.cu:622, 690-697): a different slot order leaves them where they are
    e = c1.compare(shuffled, res)
    assert all(v is True or v == 0.0 for v in e.values()), e


def _dp_flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radnerf_b200.train import GradSync
    m = _toy_model()
    sync = GradSync(m.parameters(), table_numel=1024, overlap=False)      # what GraphedTrainStep uses: no hook-driven reductions
    got = []
    for step in range(3):
        for p in m.parameters():
            if p.grad is not None:
                p.grad.zero_()
        sync.begin_step()
        _toy_loss(m, 100 * step + rank).backward()
        if step == 0:
            sync.flatten()       # after the first backward: the small gradients become views of one buffer, exchanged in place
            ptrs = {id(p): p.grad.data_ptr() for p in m.parameters() if p.grad is not None}
        sync.finish()
        got.append([None if p.grad is None else p.grad.clone() for p in m.parameters()])
        assert all(p.grad.data_ptr() == ptrs[id(p)] for p in m.parameters() if p.grad is not None)   # backward accumulated INTO the views
    flat_ok = sync.flat is not None and all(p.grad.untyped_storage().data_ptr() == sync.flat.untyped_storage().data_ptr() for p in sync.flat_params)
    out[rank] = (got, sync.bytes_last, flat_ok, [p.numel() for p in sync.flat_params])
    dist.destroy_process_group()


def test_gradsync_flat_bucket_two_ranks_gloo():
    """the in-place gradient exchange of the graphed data-parallel step: after flatten() the small parameters' .grad are views of ONE
    flat buffer that is all-reduced where it lies, the table is reduced after backward (no hooks); every step yields the mean over ranks
    of the per-rank gradients and parameters without a gradient stay without one"""
    world = 2
    out = mp.Manager().dict()
    port = 29500 + (os.getpid() + 57) % 1000
    mp.spawn(_dp_flat_worker, args=(world, port, out), nprocs=world, join=True)
    for step in range(3):
        want = None
        for r in range(world):
            m = _toy_model()
            _toy_loss(m, 100 * step + r).backward()
            g = [None if p.grad is None else p.grad / world for p in m.parameters()]
            want = g if want is None else [a if b is None else a + b for a, b in zip(want, g)]
        for r in range(world):
            for a, b in zip(out[r][0][step], want):
                assert (a is None) == (b is None)
                if a is not None:
                    assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    assert out[0][2] and out[1][2] and out[0][3] == [6, 3]
    assert out[0][1] == (1024 + 6 + 3) * 4


def test_graphed_step_capacity_has_headroom_and_hysteresis():
    """GraphedTrainStep._capacity (host logic, no GPU): 20 % headroom rounded to the bucket, kept while the estimate stays within
    [60 %, 100 %] of it, re-sized on overflow or a large shrink"""
    from radnerf_b200.train import GraphedTrainStep
    s = GraphedTrainStep.__new__(GraphedTrainStep)
    s.bucket, s.capacity, s.mc_host = 1 << 16, None, 0
    padded, cap = s._capacity(700000)
    assert padded == 700000 + (128 - 700000 % 128) and cap == 13 * 65536 and cap >= 1.2 * 700000
    s.capacity = cap
    assert s._capacity(int(0.95 * cap))[1] == cap and s._capacity(int(0.61 * cap))[1] == cap      # drift inside the band: same graph
    assert s._capacity(cap + 1)[1] > cap                                                          # overflow: grow
    assert s._capacity(int(0.3 * cap))[1] < cap                                                   # large shrink: shrink
    # top rung: with the marcher's bound known, an estimate whose capacity comes within 25 % of rays x max_steps gets the bound itself,
    # and nothing the estimate does afterwards (it cannot exceed the bound) asks for another capture
    s.capacity, s.max_samples = None, 65536 * 16
    top = ((65536 * 16 + 128 + s.bucket - 1) // s.bucket) * s.bucket
    assert s._capacity(200000)[1] < 0.75 * top                                                    # early run: a small rung
    assert s._capacity(690000)[1] == top
    s.capacity = top
    for mc in (700000, 910000, 65536 * 16, int(0.61 * top)):
        assert s._capacity(mc)[1] == top


def test_ops_frame_is_registered_by_the_oracle_only():
    """the reference's op-by-op inference loop is test infrastructure: radnerf_b200.model has no such loop of its own and raises without
    the registration that importing oracle.ops_frame (tests/conftest.py) performs"""
    from radnerf_b200 import model as M
    from oracle import ops_frame
    assert M._OPS_FRAME is ops_frame.ops_frame
    src = open(os.path.join(ROOT, "rad-nerf_b200", "radnerf_b200", "model.py")).read()
    assert "march_rays(" not in src and "composite_rays(" not in src
    saved = M._OPS_FRAME
    try:
        M.register_ops_frame(None)
        from oracle.cpu_backend import CPUOps
        import bench
        m = bench.make_model("cpu", ops=CPUOps(), fp16=False)
        z = torch.zeros(1, 4, 3)
        with pytest.raises(RuntimeError, match="test infrastructure"):
            m.render(z, z + 1, None, torch.zeros(1, 4, 2), torch.zeros(1, 6), path="ops")
    finally:
        M.register_ops_frame(saved)
