"""GPU tests of the fused / tensor-core path: the tcgen05 plumbing in isolation, then the fused frame renderer against the
op-by-op path (which is itself parity-checked against the reference kernels in test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,N", [(16, 16), (32, 64), (64, 64), (64, 16), (64, 80), (80, 64), (96, 64), (128, 128)])
def test_tcgen05_gemm_tile_matches_torch_fp32(K, N):
    """out = A @ W^T with fp16 operands and fp32 accumulation in TMEM vs a plain PyTorch fp32 matmul of the same fp16 values."""
    from radnerf_b200 import abi as L
    g = torch.Generator(device="cpu").manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).half().to(DEV)
    W = torch.randn(N, K, generator=g).half().to(DEV)
    out = torch.full((128, N), float("nan"), device=DEV)
    L.check(L.lib().rn_selftest_umma(L.ptr(A), L.ptr(W), L.ptr(out), K, N, L.cur_stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    assert torch.isfinite(out).all()
    # products of fp16 values are exact in fp32; only the summation order differs
    assert (out - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------------------------ fused frame
def _scene(hw, torso=True, trained_like=True, seed=0):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from radnerf_b200 import synthetic as syn
    model = bench.make_model(DEV, seed=seed)
    if trained_like:  # tables with O(1) entries make every stage numerically visible (random-init tables are ~1e-4)
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        for enc in (model.encoder, model.encoder_ambient, model.torso_encoder):
            with torch.no_grad():
                enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1).to(DEV))
    frames, intr, bg = bench.make_frames(hw, 4)
    out = []
    for f in frames:
        ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
        out.append(dict(ro=torch.from_numpy(ro).to(DEV)[None], rd=torch.from_numpy(rd).to(DEV)[None],
                        auds=torch.from_numpy(f["auds"]).to(DEV), pose6=torch.from_numpy(f["pose6"]).to(DEV),
                        eye=torch.from_numpy(f["eye"]).to(DEV)))
    return model, out, torch.from_numpy(bg).to(DEV)[None]


def _render(model, f, bg, path):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return model.render(f["ro"], f["rd"], f["auds"], bg, f["pose6"], eye=f["eye"], index=0, bg_color=None, perturb=False,
                            path=path, **model.opt.render_kwargs())


@pytest.mark.parametrize("hw,trained_like", [(64, True), (128, True), (128, False)])
def test_fused_frame_matches_op_by_op_path(hw, trained_like):
    """The fused renderer against the reference-ordered op-by-op path (itself bit-checked against the reference kernels).
    Ray schedule (n_alive, n_step per iteration) and sample counts must be IDENTICAL; image/depth/weights within 1e-3 (north_star's fp16 bound):
    the fused MLP accumulates in a different order than cuBLAS (fp32 accumulation of fp16 products), and every layer output
    is rounded to fp16 in both, so one fp16 ulp (~5e-4 relative) of drift per layer is the expected scale."""
    from radnerf_b200 import frame
    model, frames, bg = _scene(hw, trained_like=trained_like)
    for i, f in enumerate(frames[:3]):
        model.enc_a = None if i == 0 else enc_a_ref
        ref = _render(model, f, bg, "ops")
        enc_a_ref = model.enc_a.clone()
        sched_ref = list(model.last_frame_stats)
        model.enc_a = None if i == 0 else enc_a_fused
        out = _render(model, f, bg, "fused")
        torch.cuda.synchronize()
        enc_a_fused = model.enc_a.clone()
        sched = frame.frame_stats(model)
        assert [(a, s) for a, s, _ in sched] == [(a, s) for a, s, _ in sched_ref], (sched, sched_ref)
        # lip-smoothed audio code: fp16 conv/linear stack, fp32 attention
        assert (enc_a_fused - enc_a_ref).abs().max().item() <= 1e-3 * max(1.0, enc_a_ref.abs().max().item())
        d_img = (out["image"] - ref["image"]).abs().max().item()
        d_dep = (out["depth"] - ref["depth"]).abs().max().item()
        d_ta = (out["torso_alpha"] - ref["torso_alpha"]).abs().max().item()
        print(f"hw={hw} trained_like={trained_like} frame={i}: |d image|={d_img:.2e} |d depth|={d_dep:.2e} |d torso_alpha|={d_ta:.2e}")
        assert d_img <= 1e-3 and d_dep <= 1e-3 and d_ta <= 1e-3
        assert (out["torso_color"] - ref["torso_color"].view(-1, 3)).abs().max().item() <= 1e-3


def test_fused_frame_sample_counts_match_reference_slots():
    """n_samples per iteration == number of non-empty slots the reference-layout marcher produces."""
    import raymarching as rm
    from radnerf_b200 import frame
    model, frames, bg = _scene(96, trained_like=False)
    f = frames[0]
    _render(model, f, bg, "fused")
    sched = frame.frame_stats(model)
    ro, rd = f["ro"][0], f["rd"][0]
    nears, fars = rm.near_far_from_aabb(ro, rd, model.aabb_infer, model.min_near)
    N = ro.shape[0]
    alive = torch.arange(N, dtype=torch.int32, device=DEV)
    x, d, dl = rm.march_rays(N, 1, alive, nears.clone(), ro, rd, model.bound, model.density_bitfield, model.cascade, model.grid_size,
                             nears, fars, 128, False, model.opt.dt_gamma, model.opt.max_steps)
    assert sched[0][0] == N and sched[0][1] == 1 and sched[0][2] == int((dl[:, 0] > 0).sum().item())


def test_frame_streamer_delivers_the_frames_model_render_produces():
    """The host-facing pipeline (one H2D block, rays on the device, graph replay, double-buffered D2H) must deliver exactly
    the images a plain model.render call produces for the same host inputs, in order, while keeping two frames in flight."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from radnerf_b200.stream import FrameStreamer, pack_inputs
    from radnerf_b200.rays import RayGenerator
    hw = 64
    model = bench.make_model(DEV, seed=3)
    model.smooth_lips = False  # frames are rendered twice below: keep them independent of the call history
    frames, intr, bg = bench.make_frames(hw, 5)
    bg_t = torch.from_numpy(bg).to(DEV)
    kw = model.opt.render_kwargs()
    raygen = RayGenerator(hw, hw, intr, torch.device(DEV))
    want = []
    for f in frames:
        ro, rd = raygen(torch.from_numpy(f["pose"]).to(DEV))
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = model.render(ro[None], rd[None], torch.from_numpy(f["auds"]).to(DEV), bg_t[None], torch.from_numpy(f["pose6"]).to(DEV),
                               eye=torch.from_numpy(f["eye"]).to(DEV), index=0, bg_color=None, perturb=False, path="fused", **kw)
        want.append(out["image"].reshape(-1, 3).cpu().clone())
    streamer = FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, use_eye=True, depth=2, **kw)
    got = [img.clone() for img in streamer.render_all([pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames])]
    assert len(got) == len(want) and streamer.in_flight() == 0
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert not torch.equal(got[0], got[1])  # the frames differ, so order and slot reuse are really exercised


def test_frame_parallel_slices_reproduce_the_whole_sequence_run():
    """render_sequence(world=2) must give, for each rank's slice, the frames a single run over the whole sequence gives --
    including the lip-smoothing state that couples consecutive frames (primed over the frames in front of the slice)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from radnerf_b200.stream import pack_inputs, render_sequence, sequence_slice
    hw, n = 64, 60
    model = bench.make_model(DEV, seed=5)
    assert model.smooth_lips
    frames, intr, bg = bench.make_frames(hw, n)
    bg_t = torch.from_numpy(bg).to(DEV)
    kw = model.opt.render_kwargs()
    packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames]
    whole = {i: img.clone() for i, img in render_sequence(model, packed, hw, hw, intr, bg_t, frames[0]["auds"].shape, **kw)}
    assert sorted(whole) == list(range(n))
    assert sequence_slice(n, 2, 0) == (0, 30) and sequence_slice(n, 2, 1) == (30, 60) and sequence_slice(7, 3, 0) == (0, 3)
    for rank in (0, 1):
        part = {i: img.clone() for i, img in render_sequence(model, packed, hw, hw, intr, bg_t, frames[0]["auds"].shape, world=2, rank=rank, **kw)}
        lo, hi = sequence_slice(n, 2, rank)
        assert sorted(part) == list(range(lo, hi))
        for i in part:
            assert (part[i] - whole[i]).abs().max().item() <= 1e-6, (rank, i)
    # the smoothing state matters: without priming, the first frame of rank 1's slice would differ visibly
    model.enc_a = None
    assert not torch.equal(whole[29], whole[30])


def _custom_scene(hw, n_frames=2, seed=0, **opt_kw):
    """like _scene but for an arbitrary reference configuration (BASELINE.json configs[1] and configs[4])"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.posemath import convert_poses
    torch.manual_seed(seed)
    model = NeRFNetwork(Options(smooth_lips=True, fp16=True, exp_eye=True, **opt_kw))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    model.density_grid.copy_(torch.from_numpy(grid))
    model.mean_density = float(np.clip(grid, 0, None).mean())
    model.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(model.mean_density, model.density_thresh))))
    if model.torso:
        tg = syn.torso_density_grid(128)
        model.density_grid_torso.copy_(torch.from_numpy(tg))
        model.mean_density_torso = float(tg.mean())
    model = model.eval().to(DEV)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    encs = [model.encoder, model.encoder_ambient] + ([model.torso_encoder] if model.torso else [])
    for enc in encs:
        with torch.no_grad():
            enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1).to(DEV))
    bank = syn.audio_feature_bank(600, model.audio_in_dim, 16, seed=0)
    intr = syn.intrinsics_for(hw, hw)
    frames = []
    for i in range(n_frames):
        pose = syn.orbit_pose(yaw_deg=8.0 * i - 4.0, pitch_deg=2.0)
        ro, rd = syn.get_rays(pose, intr, hw, hw)
        frames.append(dict(ro=torch.from_numpy(ro).to(DEV)[None], rd=torch.from_numpy(rd).to(DEV)[None],
                           auds=torch.from_numpy(syn.audio_window(bank, 8 + i, model.att)).to(DEV),
                           pose6=convert_poses(torch.from_numpy(pose)[None]).to(DEV), eye=torch.tensor([[0.25]], device=DEV)))
    return model, frames, torch.from_numpy(syn.get_bg_coords(hw, hw)).to(DEV)[None]


@pytest.mark.parametrize("name,hw,opt_kw", [
    ("configs[1]: head only, 450x450, wav2vec 44-d", 450, dict(torso=False)),
    ("configs[4]: DeepSpeech 29-d with eye / individual codes, torso", 128, dict(torso=True, asr_model="deepspeech")),
    ("32-d audio features (the reference's third branch)", 96, dict(torso=True, asr_model="hubert")),
    ("configs[2] at full size: 512x512 head+torso", 512, dict(torso=True)),
])
def test_fused_frame_on_the_reference_configurations(name, hw, opt_kw):
    """BASELINE.json's other configurations as parity cases: fused frame vs the op-by-op path, identical ray schedule,
    image / depth / torso within 1e-3 (fp16 tables with O(1) entries)."""
    from radnerf_b200 import frame
    model, frames, bg = _custom_scene(hw, **opt_kw)
    assert frame.supported(model), name
    enc_ops = enc_fused = None
    for i, f in enumerate(frames):
        model.enc_a = enc_ops
        ref = _render(model, f, bg, "ops")
        enc_ops = model.enc_a.clone()
        sched_ref = [(a, s) for a, s, _ in model.last_frame_stats]
        model.enc_a = enc_fused
        out = _render(model, f, bg, "fused")
        torch.cuda.synchronize()
        enc_fused = model.enc_a.clone()
        assert [(a, s) for a, s, _ in frame.frame_stats(model)] == sched_ref, name
        assert (enc_fused - enc_ops).abs().max().item() <= 1e-3 * max(1.0, enc_ops.abs().max().item())
        assert (out["image"] - ref["image"]).abs().max().item() <= 1e-3, name
        assert (out["depth"] - ref["depth"]).abs().max().item() <= 1e-3, name
        if model.torso:
            assert (out["torso_alpha"] - ref["torso_alpha"]).abs().max().item() <= 1e-3, name


def test_pipelined_frames_keep_the_lip_smoothing_chain():
    """Three frames in flight (FramePipeline lanes) must produce the images of a strictly sequential run, including the
    lip-smoothing EMA that makes frame i depend on frames < i (conditioning kernels run in frame order on their own stream)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from radnerf_b200.stream import FrameStreamer, pack_inputs
    from radnerf_b200.rays import RayGenerator
    hw, n = 64, 9
    model = bench.make_model(DEV, seed=7)
    assert model.smooth_lips
    frames, intr, bg = bench.make_frames(hw, n)
    bg_t = torch.from_numpy(bg).to(DEV)
    kw = model.opt.render_kwargs()
    raygen = RayGenerator(hw, hw, intr, torch.device(DEV))
    model.enc_a = None
    want = []
    for f in frames:
        ro, rd = raygen(torch.from_numpy(f["pose"]).to(DEV))
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = model.render(ro[None], rd[None], torch.from_numpy(f["auds"]).to(DEV), bg_t[None], torch.from_numpy(f["pose6"]).to(DEV),
                               eye=torch.from_numpy(f["eye"]).to(DEV), index=0, bg_color=None, perturb=False, path="fused", **kw)
        want.append(out["image"].reshape(-1, 3).cpu().clone())
    model.enc_a = None
    streamer = FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, use_eye=True, depth=3, **kw)
    got = [img.clone() for img in streamer.render_all([pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames])]
    assert len(got) == n
    for i, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a, b), i
    # and the chain is really there: rendering frame 5 without its history gives a different image
    model.enc_a = None
    f = frames[5]
    ro, rd = raygen(torch.from_numpy(f["pose"]).to(DEV))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        lone = model.render(ro[None], rd[None], torch.from_numpy(f["auds"]).to(DEV), bg_t[None], torch.from_numpy(f["pose6"]).to(DEV),
                            eye=torch.from_numpy(f["eye"]).to(DEV), index=0, bg_color=None, perturb=False, path="fused", **kw)
    assert not torch.equal(lone["image"].reshape(-1, 3).cpu(), want[5])


def test_uint8_output_stage_equals_the_reference_host_expression():
    """FrameStreamer(output="uint8") must deliver (pred * 255).astype(np.uint8) of the fp32 frame (nerf/utils.py:952-960), through
    both the first-frame Python path and the one-call C path"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from radnerf_b200.stream import FrameStreamer, pack_inputs
    hw = 64
    model = bench.make_model(DEV, seed=11)
    frames, intr, bg = bench.make_frames(hw, 7)
    bg_t = torch.from_numpy(bg).to(DEV)
    kw = model.opt.render_kwargs()
    packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames]
    model.enc_a = None
    f32 = [img.clone().numpy() for img in FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, depth=2, **kw).render_all(packed)]
    model.enc_a = None
    u8 = [img.clone().numpy() for img in FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, depth=2, output="uint8", **kw).render_all(packed)]
    assert len(u8) == len(f32) == 7
    for a, b in zip(u8, f32):
        assert a.dtype == np.uint8 and np.array_equal(a, (b * 255).astype(np.uint8))
    assert u8[0].min() < 255   # not just background
