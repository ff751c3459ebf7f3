"""SURVEY 8(a) rows at BASELINE sizes: every reference kernel next to its replacement ON THE SAME INPUTS on the same GPU --
parity (bit-exact where the contract says so) and CUDA-event timings.  The reference side is the reference's own
extensions compiled by oracle/build_ref.py (oracle/_ref/*.so); without them the comparison half is skipped and only the
full-size self-consistency properties run.  Writes the table to gpurun_out/kernel_rows.json (copied to profiles/)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TABLE = {}


def _ref(name):
    from oracle import ref_backend
    if not ref_backend.available():
        pytest.skip("oracle/_ref/*.so not built (python oracle/build_ref.py)")
    return ref_backend.backend(name)


def _time(fn, reps=15):
    """device time of ONE call: a spin kernel keeps the GPU busy while the host enqueues (event, call, event), so the
    Python / ctypes / pybind call overhead (5-15 us, more than several of these kernels take) is not part of the number"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        torch.cuda._sleep(400000)   # ~0.2 ms
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def _row(key, ours_ms, ref_ms, units, bytes_per_unit, note=""):
    from radnerf_b200 import roofline
    hbm = roofline.peaks()[0]
    TABLE[key] = {"ours_ms": ours_ms, "reference_ms": ref_ms, "speedup": ref_ms / ours_ms, "units": units,
                  "algorithmic_bytes_per_unit": bytes_per_unit, "ours_GBps": units * bytes_per_unit / ours_ms / 1e6,
                  "ours_frac_of_hbm_peak": units * bytes_per_unit / ours_ms / 1e6 / hbm, "note": note}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(TABLE, open(os.path.join(ROOT, "gpurun_out", "kernel_rows.json"), "w"), indent=1)
    print(f"{key}: ours {ours_ms:.4f} ms, reference {ref_ms:.4f} ms ({ref_ms / ours_ms:.2f}x), "
          f"{TABLE[key]['ours_GBps']:.0f} GB/s = {100 * TABLE[key]['ours_frac_of_hbm_peak']:.1f}% of HBM peak")


@pytest.fixture(scope="module")
def scene():
    import bench
    from radnerf_b200 import roofline, synthetic as syn
    model = bench.make_model(DEV)
    g = torch.Generator(device="cpu").manual_seed(1)
    for enc in (model.encoder, model.encoder_ambient):
        with torch.no_grad():
            enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1).to(DEV))
    frames, intr, bg = bench.make_frames(512, 1)
    ro, rd = syn.get_rays(frames[0]["pose"], intr, 512, 512)
    f = dict(ro=torch.from_numpy(ro).to(DEV)[None], rd=torch.from_numpy(rd).to(DEV)[None])
    kw = model.opt.render_kwargs()
    x, dirs, nears, fars = roofline.frame_samples(model, f, kw)   # the occupied samples of a 512x512 frame, ~0.8 M
    return dict(model=model, ro=f["ro"][0].contiguous(), rd=f["rd"][0].contiguous(), x=x, dirs=dirs.contiguous(), nears=nears, fars=fars, kw=kw)


def _grid_args(enc):
    return (enc.input_dim, enc.level_dim, enc.num_levels, float(np.log2(enc.per_level_scale)), int(enc.base_resolution),
            enc.gridtype_id, int(enc.align_corners), enc.interp_id)


@pytest.mark.parametrize("which", ["3d", "2d"])
def test_grid_encode_forward_backward_fp16(scene, which):
    """a2/a3: kernel_grid / kernel_grid_backward vs grid_forward_kernel / grid_backward_kernel, fp16 tables (autocast path)"""
    from radnerf_b200 import abi as A
    G = _ref("_gridencoder")
    m = scene["model"]
    enc = m.encoder if which == "3d" else m.encoder_ambient
    D, Cc, L, S, H, gt, al, ip = _grid_args(enc)
    x = ((scene["x"] + m.bound) / (2 * m.bound)).contiguous() if which == "3d" else torch.rand(scene["x"].shape[0], 2, device=DEV)
    B = x.shape[0]
    table = enc.embeddings.detach().half().contiguous()
    offs = enc.offsets
    out_ref = torch.empty(L, B, Cc, device=DEV, dtype=torch.half)
    out = torch.empty(B, L * Cc, device=DEV, dtype=torch.half)
    ref_fwd = lambda: G.grid_encode_forward(x, table, offs, out_ref, B, D, Cc, L, S, H, None, gt, al, ip)
    our_fwd = lambda: A.check(A.lib().rn_grid_encode_forward(A.ptr(x), A.ptr(table), A.ptr(offs), A.ptr(out), B, D, Cc, L, S, H, None, gt, al,
                                                            ip, 1, 1, A.cur_stream()))
    ref_fwd(); our_fwd(); torch.cuda.synchronize()
    assert torch.equal(out, out_ref.permute(1, 0, 2).reshape(B, L * Cc)), "fp16 forward must be bit-identical to the reference kernel"
    bpu = 4 * D + L * (2 ** D) * Cc * 2 + L * Cc * 2
    # the reference wrapper also pays a permute copy of the [L,B,C] result (grid.py:57); kernel-only times here
    _row(f"G1 grid forward {which} fp16", _time(our_fwd), _time(ref_fwd), B, bpu)

    grad = torch.randn(B, L * Cc, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3)).half()
    grad_lbc = grad.view(B, L, Cc).permute(1, 0, 2).contiguous()
    ge_ref = torch.zeros_like(table)
    ge = torch.zeros(table.shape, device=DEV, dtype=torch.float32)
    ref_bwd = lambda: G.grid_encode_backward(grad_lbc, x, table, offs, ge_ref, B, D, Cc, L, S, H, None, None, gt, al, ip)
    our_bwd = lambda: A.check(A.lib().rn_grid_encode_backward(A.ptr(grad), A.ptr(x), A.ptr(table), A.ptr(offs), A.ptr(ge), B, D, Cc, L, S, H,
                                                             None, None, gt, al, ip, 1, 1, 0, A.cur_stream()))
    ref_bwd(); our_bwd(); torch.cuda.synchronize()
    # The corner weights of a sample sum to 1, so the sum of a level's gradient rows must equal the sum of the incoming
    # gradient columns of that level: an exact, size-independent property -- ours (fp32 accumulation) meets it to 1e-4.
    # The reference accumulates with fp16 atomics: at this size a coarse-level row receives ~1000 adds, each rounded at
    # the row's running magnitude in arrival order, so it is only good to a few percent of the level's gradient norm
    # (exact backward parity is pinned on fp32 tables by the goldens, tests/test_gpu_parity.py).
    o = offs.cpu().numpy()
    gsum = grad.view(B, L, Cc).double().sum(0)
    for l in (0, 3, 8, 15):
        a, b = ge[o[l]:o[l + 1]].double(), ge_ref[o[l]:o[l + 1]].double()
        scale = a.norm().item()
        assert (a.sum(0) - gsum[l]).abs().max().item() <= 1e-4 * max(scale, gsum[l].abs().max().item()), (l, a.sum(0), gsum[l])
        assert (a - b).norm().item() <= 0.1 * scale, (l, (a - b).norm().item(), scale)
    bpu_b = 4 * D + L * Cc * 2 + 2 * L * (2 ** D) * Cc * 2
    _row(f"G2 grid backward {which} fp16", _time(our_bwd), _time(ref_bwd), B, bpu_b, "reference: fp16 atomics into a half table; ours: fp32 vector reds")


def test_near_far_march_composite_inference(scene):
    """a5/a11/a12: near_far_from_aabb, march_rays (first iteration of a 512x512 frame), composite_rays -- bit-exact"""
    from radnerf_b200 import abi as A
    R = _ref("_raymarching_face")
    m, ro, rd = scene["model"], scene["ro"], scene["rd"]
    N = ro.shape[0]
    L_ = A.lib()
    n1, f1, n2, f2 = (torch.empty(N, device=DEV) for _ in range(4))
    ref_nf = lambda: R.near_far_from_aabb(ro, rd, m.aabb_infer, N, m.min_near, n1, f1)
    our_nf = lambda: A.check(L_.rn_near_far_from_aabb(A.ptr(ro), A.ptr(rd), A.ptr(m.aabb_infer), N, m.min_near, A.ptr(n2), A.ptr(f2), A.cur_stream()))
    ref_nf(); our_nf(); torch.cuda.synchronize()
    assert torch.equal(n1, n2) and torch.equal(f1, f2)
    _row("R1 near_far_from_aabb", _time(our_nf), _time(ref_nf), N, 32)

    n_step, max_steps, dtg = 2, scene["kw"]["max_steps"], scene["kw"]["dt_gamma"]
    alive = torch.arange(N, dtype=torch.int32, device=DEV)
    M = N * n_step + 128 - (N * n_step) % 128
    bufs = [[torch.zeros(M, 3, device=DEV), torch.zeros(M, 3, device=DEV), torch.zeros(M, 2, device=DEV)] for _ in range(2)]
    noises = torch.zeros(N, device=DEV)

    def ref_march():
        x, d, dl = bufs[0]
        R.march_rays(N, n_step, alive, n1.clone(), ro, rd, m.bound, dtg, max_steps, m.cascade, m.grid_size, m.density_bitfield, n1, f1, x, d, dl, noises)

    def our_march():
        x, d, dl = bufs[1]
        t = n1.clone()
        A.check(L_.rn_march_rays(N, n_step, A.ptr(alive), A.ptr(t), A.ptr(ro), A.ptr(rd), m.bound, dtg, max_steps, m.cascade, m.grid_size,
                                 A.ptr(m.density_bitfield), A.ptr(n1), A.ptr(f1), A.ptr(x), A.ptr(d), A.ptr(dl), A.ptr(noises), A.cur_stream()))
    ref_march(); our_march(); torch.cuda.synchronize()
    for a, b in zip(bufs[0], bufs[1]):
        assert torch.equal(a, b), "inference march outputs must be bit-identical"
    n_samples = int((bufs[0][2][:, 0] > 0).sum())
    _row("R11 march_rays (n_step=2, 512x512)", _time(our_march), _time(ref_march), N, 44 + 32 * n_samples / N, "includes the rays_t clone in both")

    sig = torch.rand(M, device=DEV) * 5
    rgb = torch.rand(M, 3, device=DEV)
    st = [[torch.zeros(N, device=DEV), torch.zeros(N, device=DEV), torch.zeros(N, 3, device=DEV)] for _ in range(2)]
    al = [alive.clone(), alive.clone()]
    rt = [n1.clone(), n1.clone()]
    ref_c = lambda: R.composite_rays(N, n_step, 1e-2, al[0], rt[0], sig, rgb, bufs[0][2], *st[0])
    our_c = lambda: A.check(L_.rn_composite_rays(N, n_step, 1e-2, A.ptr(al[1]), A.ptr(rt[1]), A.ptr(sig), A.ptr(rgb), A.ptr(bufs[1][2]),
                                                 A.ptr(st[1][0]), A.ptr(st[1][1]), A.ptr(st[1][2]), A.cur_stream()))
    ref_c(); our_c(); torch.cuda.synchronize()
    assert torch.equal(al[0], al[1]) and torch.equal(rt[0], rt[1])
    for a, b in zip(st[0], st[1]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    _row("R12 composite_rays", _time(our_c), _time(ref_c), N, 56 + 24 * n_step, "timed on an already-terminated alive list after the first call")


def test_training_march_and_composite(scene):
    """a9/a10: march_rays_train (2^16 rays), composite_rays_train forward/backward"""
    from radnerf_b200 import abi as A
    R = _ref("_raymarching_face")
    m = scene["model"]
    L_ = A.lib()
    N = 1 << 16
    sel = torch.randperm(scene["ro"].shape[0], generator=torch.Generator().manual_seed(2))[:N].to(DEV)
    ro, rd = scene["ro"][sel].contiguous(), scene["rd"][sel].contiguous()
    nears, fars = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    R.near_far_from_aabb(ro, rd, m.aabb_train, N, m.min_near, nears, fars)
    M, max_steps, dtg = N * 16, 16, 1 / 256
    noises = torch.rand(N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(4))
    outs = [[torch.zeros(M, 3, device=DEV), torch.zeros(M, 3, device=DEV), torch.zeros(M, 2, device=DEV),
             torch.empty(N, 3, dtype=torch.int32, device=DEV), torch.zeros(2, dtype=torch.int32, device=DEV)] for _ in range(2)]

    def ref_m():
        x, d, dl, rays, cnt = outs[0]
        cnt.zero_()
        R.march_rays_train(ro, rd, m.density_bitfield, m.bound, dtg, max_steps, N, m.cascade, m.grid_size, M, nears, fars, x, d, dl, rays, cnt, noises)

    def our_m():
        x, d, dl, rays, cnt = outs[1]
        cnt.zero_()
        A.check(L_.rn_march_rays_train(A.ptr(ro), A.ptr(rd), A.ptr(m.density_bitfield), m.bound, dtg, max_steps, N, m.cascade, m.grid_size, M,
                                       A.ptr(nears), A.ptr(fars), A.ptr(x), A.ptr(d), A.ptr(dl), A.ptr(rays), A.ptr(cnt), A.ptr(noises), A.cur_stream()))
    ref_m(); our_m(); torch.cuda.synchronize()
    assert torch.equal(outs[0][4], outs[1][4]), "sample / ray counters must match exactly"
    # per-ray sample counts, independent of slot order
    def counts(rays):
        c = torch.zeros(N, dtype=torch.int64, device=DEV)
        c[rays[:, 0].long()] = rays[:, 2].long()
        return c
    assert torch.equal(counts(outs[0][3]), counts(outs[1][3]))
    n_samples = int(outs[1][4][0])
    _row("R7 march_rays_train (2^16 rays)", _time(our_m), _time(ref_m), N, 48 + 32 * n_samples / N)

    x, d, dl, rays, cnt = outs[1]
    Ms = n_samples
    sig, rgb, amb = torch.rand(Ms, device=DEV) * 5, torch.rand(Ms, 3, device=DEV), torch.rand(Ms, device=DEV)
    res = [[torch.empty(N, device=DEV), torch.empty(N, device=DEV), torch.empty(N, device=DEV), torch.empty(N, 3, device=DEV)] for _ in range(2)]
    ref_f = lambda: R.composite_rays_train_forward(sig, rgb, amb, dl, rays, Ms, N, 1e-4, *res[0])
    our_f = lambda: A.check(L_.rn_composite_rays_train_forward(A.ptr(sig), A.ptr(rgb), A.ptr(amb), A.ptr(dl), A.ptr(rays), Ms, N, 1e-4,
                                                               *[A.ptr(t) for t in res[1]], A.cur_stream()))
    ref_f(); our_f(); torch.cuda.synchronize()
    for a, b in zip(res[0], res[1]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    _row("R9 composite_rays_train forward", _time(our_f), _time(ref_f), Ms, 28 + 36 * N / Ms)
    gws, gam, gim = torch.randn(N, device=DEV), torch.randn(N, device=DEV), torch.randn(N, 3, device=DEV)
    gr = [[torch.zeros(Ms, device=DEV), torch.zeros(Ms, 3, device=DEV), torch.zeros(Ms, device=DEV)] for _ in range(2)]
    ref_b = lambda: R.composite_rays_train_backward(gws, gam, gim, sig, rgb, amb, dl, rays, res[0][0], res[0][1], res[0][3], Ms, N, 1e-4, *gr[0])
    our_b = lambda: A.check(L_.rn_composite_rays_train_backward(A.ptr(gws), A.ptr(gam), A.ptr(gim), A.ptr(sig), A.ptr(rgb), A.ptr(amb), A.ptr(dl),
                                                                A.ptr(rays), A.ptr(res[1][0]), A.ptr(res[1][1]), A.ptr(res[1][3]), Ms, N, 1e-4,
                                                                *[A.ptr(t) for t in gr[1]], A.cur_stream()))
    ref_b(); our_b(); torch.cuda.synchronize()
    for a, b in zip(gr[0], gr[1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
    _row("R10 composite_rays_train backward", _time(our_b), _time(ref_b), Ms, 44 + 48 * N / Ms)


def test_occupancy_utilities_full_grid(scene):
    """a6/a7/a8: morton3D, packbits, morton3D_dilation on the full 128^3 grid -- bit-exact"""
    from radnerf_b200 import abi as A
    R = _ref("_raymarching_face")
    L_ = A.lib()
    H = 128
    n = H ** 3
    g = torch.Generator(device=DEV).manual_seed(7)
    coords = torch.randint(0, H, (n, 3), device=DEV, generator=g, dtype=torch.int32)
    i1, i2 = torch.empty(n, dtype=torch.int32, device=DEV), torch.empty(n, dtype=torch.int32, device=DEV)
    ref_m = lambda: R.morton3D(coords, n, i1)
    our_m = lambda: A.check(L_.rn_morton3D(A.ptr(coords), n, A.ptr(i2), A.cur_stream()))
    ref_m(); our_m(); torch.cuda.synchronize()
    assert torch.equal(i1, i2)
    _row("R3 morton3D (128^3 points)", _time(our_m), _time(ref_m), n, 16)
    grid = torch.rand(1, n, device=DEV, generator=g) * 20
    b1, b2 = torch.empty(n // 8, dtype=torch.uint8, device=DEV), torch.empty(n // 8, dtype=torch.uint8, device=DEV)
    ref_p = lambda: R.packbits(grid, n // 8, 10.0, b1)
    our_p = lambda: A.check(L_.rn_packbits(A.ptr(grid), n // 8, 10.0, A.ptr(b2), A.cur_stream()))
    ref_p(); our_p(); torch.cuda.synchronize()
    assert torch.equal(b1, b2)
    _row("R5 packbits (128^3)", _time(our_p), _time(ref_p), n, 4.125)
    d1, d2 = torch.empty_like(grid), torch.empty_like(grid)
    ref_d = lambda: R.morton3D_dilation(grid, 1, H, d1)
    our_d = lambda: A.check(L_.rn_morton3D_dilation(A.ptr(grid), 1, H, A.ptr(d2), A.cur_stream()))
    ref_d(); our_d(); torch.cuda.synchronize()
    assert torch.equal(d1, d2)
    _row("R6 morton3D_dilation (128^3)", _time(our_d), _time(ref_d), n, 8)


def test_freq_and_sh_encoders(scene):
    """a13/a14: freq_encode (torso coordinates, degree 10) and sh_encode (degree 4) forward"""
    from radnerf_b200 import abi as A
    F, S = _ref("_freqencoder"), _ref("_shencoder")
    L_ = A.lib()
    P = 87381
    xy = torch.rand(P, 2, device=DEV) * 2 - 1
    o1, o2 = torch.empty(P, 42, device=DEV), torch.empty(P, 42, device=DEV)
    ref_f = lambda: F.freq_encode_forward(xy, P, 2, 10, 42, o1)
    our_f = lambda: A.check(L_.rn_freq_encode_forward(A.ptr(xy), P, 2, 10, 42, A.ptr(o2), A.cur_stream()))
    ref_f(); our_f(); torch.cuda.synchronize()
    assert torch.allclose(o1, o2, rtol=0, atol=1e-5)
    _row("F1 freq_encode forward (2 -> 42)", _time(our_f), _time(ref_f), P, 176)
    dirs = scene["dirs"]
    M = dirs.shape[0]
    s1, s2 = torch.empty(M, 16, device=DEV), torch.empty(M, 16, device=DEV)
    ref_s = lambda: S.sh_encode_forward(dirs, s1, M, 3, 4, None)
    our_s = lambda: A.check(L_.rn_sh_encode_forward(A.ptr(dirs), A.ptr(s2), M, 3, 4, None, A.cur_stream()))
    ref_s(); our_s(); torch.cuda.synchronize()
    assert torch.allclose(s1, s2, rtol=1e-5, atol=1e-6)
    _row("S1 sh_encode forward (degree 4)", _time(our_s), _time(ref_s), M, 76)
