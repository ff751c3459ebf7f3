"""GPU parity tests (-m gpu): every operator goes through the reference-shaped Python API -> C ABI -> sm_100a kernels and
is compared (a) with the golden vectors produced by the reference's own CUDA kernels (tests/golden/) and (b) with the
CPU oracle on the same seeded inputs.

Tolerances (north_star): bit-exact for Morton codes, bitfields, sample counts and every marcher output; <= 1e-5
max-abs for fp32 encodings / composited values / gradients, <= 1e-3 for fp16 tables.  Where a looser bound is used the
reason is stated at the assertion.
"""
import numpy as np
import pytest
import torch

import golden_cases as gc
from conftest import golden

pytestmark = pytest.mark.gpu

DEV = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N_(t):
    return t.detach().cpu().numpy()


def maxabs(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0


@pytest.fixture(scope="module")
def abi():
    from radnerf_b200 import abi as L
    L.lib()
    return L


def device_scales(abi, S, H, L):
    sc = torch.empty(L, device=DEV)
    abi.check(abi.lib().rn_grid_level_geometry(float(S), H, L, abi.ptr(sc), None, abi.cur_stream()))
    return N_(sc)


# ------------------------------------------------------------------------------------------------ grid encoder
def run_grid_fwd(abi, c, layout=1):
    x, table, offs = T(c["inputs"]), T(c["table"]), T(c["offsets"])
    B, D, C, L, H = c["B"], c["D"], c["C"], c["L"], c["H"]
    S = float(np.log2(c["per_level_scale"]))
    out = torch.empty((B, L * C) if layout == 1 else (L, B, C), device=DEV, dtype=table.dtype)
    dy = torch.empty(B, L * D * C, device=DEV, dtype=table.dtype) if c["dy"] else None
    abi.check(abi.lib().rn_grid_encode_forward(abi.ptr(x), abi.ptr(table), abi.ptr(offs), abi.ptr(out), B, D, C, L, S, H,
                                               abi.ptr(dy), c["gridtype"], int(c["align"]), c["interp"], int(c["half"]),
                                               layout, abi.cur_stream()))
    return out, dy


@pytest.mark.parametrize("name", list(gc.GRID_CASES))
def test_grid_forward_vs_reference_golden_and_oracle(abi, oracle, name):
    c = gc.grid_case(name)
    g = golden("grid_" + name)
    out, dy = run_grid_fwd(abi, c)
    tol = 1e-3 if c["half"] else 1e-5
    # (a) the reference's CUDA kernel.  fp16 tables: bit-identical (we reproduce c10::Half's per-corner rounding);
    #     fp32: bit-identical too as long as the FMA placement matches -- asserted exactly.
    assert np.array_equal(N_(out), g["out"]), f"max diff {maxabs(N_(out).astype(np.float32), g['out'].astype(np.float32))}"
    if c["dy"]:
        assert maxabs(N_(dy).astype(np.float32), g["dy_dx"].astype(np.float32)) <= tol * max(1.0, float(np.abs(g["dy_dx"].astype(np.float32)).max()))
    # (b) the CPU oracle with the level scales the device computed
    sc = device_scales(abi, np.log2(c["per_level_scale"]), c["H"], c["L"])
    assert np.array_equal(sc, g["scales"])
    o_out, o_dy = oracle.grid_encode_forward(c["inputs"], c["table"], c["offsets"], c["per_level_scale"], c["H"], c["dy"],
                                             c["gridtype"], c["align"], c["interp"], scales=sc)
    assert np.array_equal(N_(out), o_out)
    if c["dy"]:
        assert maxabs(N_(dy).astype(np.float32), o_dy.astype(np.float32)) <= tol * max(1.0, float(np.abs(o_dy.astype(np.float32)).max()))
    # the reference's own [L,B,C] layout is available too and holds the same numbers
    out_lbc, _ = run_grid_fwd(abi, c, layout=0)
    assert np.array_equal(N_(out_lbc.permute(1, 0, 2).reshape(c["B"], -1)), N_(out))
    # out-of-range coordinates give zeros (gridencoder.cu:110-135)
    assert not N_(out)[:2].any()


@pytest.mark.parametrize("name", list(gc.GRID_CASES))
def test_grid_backward_vs_reference_golden_and_oracle(abi, oracle, name):
    c = gc.grid_case(name)
    g = golden("grid_" + name)
    x, table, offs, grad = T(c["inputs"]), T(c["table"]), T(c["offsets"]), T(c["grad"])
    B, D, C, L, H = c["B"], c["D"], c["C"], c["L"], c["H"]
    S = float(np.log2(c["per_level_scale"]))
    _, dy = run_grid_fwd(abi, c)
    ge = torch.zeros(table.shape, device=DEV, dtype=torch.float32)
    gi = torch.empty(B, D, device=DEV, dtype=table.dtype) if c["dy"] else None
    abi.check(abi.lib().rn_grid_encode_backward(abi.ptr(grad), abi.ptr(x), abi.ptr(table), abi.ptr(offs), abi.ptr(ge), B, D,
                                                C, L, S, H, abi.ptr(dy), abi.ptr(gi), c["gridtype"], int(c["align"]),
                                                c["interp"], int(c["half"]), 1, 0, abi.cur_stream()))
    ge = N_(ge).astype(np.float64)
    sc = device_scales(abi, np.log2(c["per_level_scale"]), H, L)
    o_ge, o_gi = oracle.grid_encode_backward(c["grad"], c["inputs"], c["offsets"], c["per_level_scale"], H, table.shape[0], C,
                                             dy_dx=N_(dy) if c["dy"] else None, gridtype=c["gridtype"],
                                             align_corners=c["align"], interpolation=c["interp"], scales=sc)
    scale = max(1.0, float(np.abs(o_ge).max()))
    # oracle (double accumulation).  fp16: the oracle rounds each contribution to half like the reference, we keep fp32.
    assert maxabs(ge, o_ge) <= (1e-3 if c["half"] else 1e-5) * scale
    # reference CUDA: its fp16 path ACCUMULATES in half with atomics (each add rounds to 11 bits, order arbitrary), so the
    # comparison bound is a few half-ulps of the largest sum; its fp32 path only differs by summation order.
    ref_tol = (4e-3 if c["half"] else 2e-5) * scale
    assert maxabs(ge[c["bwd_rows"]], g["grad_emb_rows"]) <= ref_tol
    lv = np.stack([ge[c["offsets"][l]:c["offsets"][l + 1]].sum(0) for l in range(L)])
    assert maxabs(lv, g["grad_emb_level_sums"]) <= (0.05 if c["half"] else 1e-3) * max(1.0, float(np.abs(lv).max()))
    if c["dy"]:
        gs = max(1.0, float(np.abs(o_gi).max()))
        assert maxabs(N_(gi).astype(np.float64), o_gi) <= (2e-3 if c["half"] else 1e-5) * gs
        # reference accumulates the input gradient in scalar_t (half: running rounding) -> looser bound for fp16
        assert maxabs(N_(gi).astype(np.float64), g["grad_inputs"]) <= (2e-2 if c["half"] else 2e-5) * gs
    # the fp16 atomics target (reference behaviour) is still available through the ABI
    if c["half"] and C % 2 == 0:
        ge16 = torch.zeros(table.shape, device=DEV, dtype=torch.float16)
        abi.check(abi.lib().rn_grid_encode_backward(abi.ptr(grad), abi.ptr(x), abi.ptr(table), abi.ptr(offs), abi.ptr(ge16), B,
                                                    D, C, L, S, H, None, None, c["gridtype"], int(c["align"]), c["interp"], 1,
                                                    1, 1, abi.cur_stream()))
        assert maxabs(N_(ge16.float()), o_ge) <= 4e-3 * scale


@pytest.mark.parametrize("half", [False, True])
def test_grid_backward_2d_clustered_inputs_warp_aggregation(abi, oracle, half):
    """the ambient grid's real input: coordinates clustered around one point (plus exact duplicates, a few out-of-range
    rows and a ragged tail), so that the lanes of a warp share cells on the coarse levels and the warp-aggregated path of
    grid_backward_shared_cell_kernel runs; uniform rows mixed in keep the per-lane path in the same launch.
    Oracle: double accumulation.  Tolerance: fp32 summation order only (1e-5 of the largest sum; 1e-3 for fp16 gradients,
    where the oracle rounds each contribution to half and we keep fp32)"""
    rng = np.random.default_rng(21)
    B, D, C, L, H = 20011, 2, 2, 16, 16
    offs, pls = oracle.grid_offsets(D, L, C, H, 16, desired_resolution=2048)     # the ambient grid of nerf/network.py
    x = (0.52 + 0.004 * rng.standard_normal((B, D))).astype(np.float32)
    x[1000:1400] = x[1000]                                   # exact duplicates: whole warps on one point
    x[5000:7000] = rng.random((2000, D), dtype=np.float32)   # uniform rows: every lane its own cell
    x[::977] = 1.5                                           # out of range: no table contribution
    grad = rng.standard_normal((B, L * C)).astype(np.float16 if half else np.float32)
    n_rows = int(offs[-1])
    xt, ot, gt_ = T(x), T(offs), T(grad)
    table = torch.zeros(n_rows, C, device=DEV, dtype=torch.float16 if half else torch.float32)
    ge = torch.zeros(n_rows, C, device=DEV, dtype=torch.float32)
    S = float(np.log2(pls))
    abi.check(abi.lib().rn_grid_encode_backward(abi.ptr(gt_), abi.ptr(xt), abi.ptr(table), abi.ptr(ot), abi.ptr(ge), B, D, C, L, S, H,
                                                None, None, 0, 0, 0, int(half), 1, 0, abi.cur_stream()))
    sc = device_scales(abi, np.log2(pls), H, L)
    o_ge, _ = oracle.grid_encode_backward(grad, x, offs, pls, H, n_rows, C, scales=sc)
    scale = float(np.abs(o_ge).max())
    assert scale > 50.0                                      # thousands of samples really pile onto single rows
    assert maxabs(N_(ge).astype(np.float64), o_ge) <= (1e-3 if half else 1e-5) * scale
    # and through the reference's fp16 accumulation target.  Every atomic add rounds the running sum to 11 bits, and here
    # hundreds to thousands of adds land on one row whose sum reaches ~100 (half ulp 0.06): the target itself is only good
    # to a few percent of the largest sum (the reference's own kernel has the same property); fewer adds -- what the warp
    # aggregation gives -- can only make it better
    if half:
        ge16 = torch.zeros(n_rows, C, device=DEV, dtype=torch.float16)
        abi.check(abi.lib().rn_grid_encode_backward(abi.ptr(gt_), abi.ptr(xt), abi.ptr(table), abi.ptr(ot), abi.ptr(ge16), B, D, C, L, S,
                                                    H, None, None, 0, 0, 0, 1, 1, 1, abi.cur_stream()))
        assert maxabs(N_(ge16.float()).astype(np.float64), o_ge) <= 5e-2 * scale


def test_grid_tv_gradient(abi, oracle):
    for name in ("g2_f32_dy", "g3_hash_sm_f32", "g2_align_c1"):
        c = gc.grid_case(name)
        g = golden("grid_" + name)
        x, table, offs = T(c["inputs"]), T(c["table"]), T(c["offsets"])
        grad = torch.zeros_like(table)
        abi.check(abi.lib().rn_grad_total_variation(abi.ptr(x), abi.ptr(table), abi.ptr(grad), abi.ptr(offs), 1e-3, c["B"],
                                                    c["D"], c["C"], c["L"], float(np.log2(c["per_level_scale"])), c["H"],
                                                    c["gridtype"], int(c["align"]), 0, abi.cur_stream()))
        sc = device_scales(abi, np.log2(c["per_level_scale"]), c["H"], c["L"])
        o = oracle.grad_total_variation(c["inputs"], c["table"], c["offsets"], c["per_level_scale"], c["H"], 1e-3,
                                        c["gridtype"], c["align"], scales=sc)
        assert maxabs(N_(grad), o) <= 1e-5
        assert maxabs(N_(grad)[c["bwd_rows"]], g["tv_rows"]) <= 1e-5


def test_grid_encoder_module_contract():
    """GridEncoder keeps the reference's attributes, state-dict keys, level table and AMP behaviour."""
    from gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16,
                      desired_resolution=2048, gridtype='tiled').to(DEV)
    assert list(enc.state_dict().keys()) == ["embeddings", "offsets"]
    assert tuple(enc.embeddings.shape) == (903480, 2) and enc.offsets.dtype == torch.int32
    assert enc.offsets.tolist()[:6] == [0, 4920, 18744, 51512, 117048, 182584] and enc.output_dim == 32
    assert abs(enc.per_level_scale - 1.381912879967776) < 1e-12
    x = (torch.rand(1000, 3, device=DEV) * 2 - 1)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    y32 = enc(x)
    assert y32.dtype == torch.float32 and y32.shape == (1000, 32)
    with torch.autocast("cuda", dtype=torch.float16):
        y16 = enc(x)
    assert y16.dtype == torch.float16
    assert (y16.float() - y32).abs().max().item() < 5e-3
    # gradients: table grad arrives in the parameter's dtype; input grad only when inputs require it
    xg = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = enc(xg)
    (out.float() ** 2).sum().backward()
    assert enc.embeddings.grad.dtype == torch.float32 and enc.embeddings.grad.abs().sum() > 0
    assert xg.grad is not None and xg.grad.shape == x.shape and xg.grad.dtype == torch.float32
    # prefix shapes are preserved
    assert enc(x.view(10, 100, 3)).shape == (10, 100, 32)
    with pytest.raises(RuntimeError):
        GridEncoder(input_dim=3, level_dim=3).to(DEV)(x)  # C must be 1, 2, 4 or 8


def test_grid_backward_matches_autograd_of_a_torch_fp32_restatement():
    """fp32 kernel vs plain PyTorch (dense bilinear lookup written with torch ops) for a 2-D dense level set."""
    from gridencoder import GridEncoder
    enc = GridEncoder(input_dim=2, num_levels=4, level_dim=2, base_resolution=8, log2_hashmap_size=14,
                      per_level_scale=2.0, gridtype='tiled').to(DEV)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    x = (torch.rand(4096, 2, device=DEV) * 2 - 1).requires_grad_(True)
    y = enc(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    g_table, g_x = enc.embeddings.grad.clone(), x.grad.clone()

    tab = enc.embeddings.detach().clone().requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    u = (x2 + 1) / 2
    outs = []
    for l in range(4):
        scale = 2.0 ** l * 8 - 1.0
        res = int(np.ceil(scale)) + 1
        pos = u * scale + 0.5
        p0 = pos.floor()
        f = pos - p0
        p0 = p0.long()
        acc = 0
        for dx in (0, 1):
            for dy in (0, 1):
                wgt = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1])
                idx = (p0[:, 0] + dx) + (p0[:, 1] + dy) * (res + 1) + int(enc.offsets[l])
                acc = acc + wgt[:, None] * tab[idx]
        outs.append(acc)
    y_ref = torch.cat(outs, 1)
    assert (y_ref - y).abs().max().item() < 1e-5
    (y_ref * w).sum().backward()
    assert (tab.grad - g_table).abs().max().item() < 1e-4 * max(1.0, tab.grad.abs().max().item())
    assert (x2.grad - g_x).abs().max().item() < 1e-3 * max(1.0, x2.grad.abs().max().item())


# ------------------------------------------------------------------------------------------------ utilities
def test_integer_utilities_bit_exact(oracle):
    import raymarching as rm
    c, g = gc.util_case(), golden("utils")
    m = N_(rm.morton3D(T(c["coords"])))
    assert m.dtype == np.int32 and np.array_equal(m, g["morton"]) and np.array_equal(m, oracle.morton3D(c["coords"]))
    assert m[:4].tolist() == [1, 2, 4, 2097151]
    inv = N_(rm.morton3D_invert(T(c["indices"])))
    assert np.array_equal(inv, g["invert"]) and np.array_equal(inv, oracle.morton3D_invert(c["indices"]))
    assert np.array_equal(N_(rm.morton3D_invert(rm.morton3D(T(c["coords"])))), c["coords"])  # invert o morton = id
    bits = N_(rm.packbits(T(c["grid32"]), c["thresh"]))
    assert bits.dtype == np.uint8 and np.array_equal(bits, g["bits32"]) and np.array_equal(bits, oracle.packbits(c["grid32"], c["thresh"]))
    assert np.array_equal(bits, np.packbits((c["grid32"].reshape(-1) > np.float32(c["thresh"])), bitorder="little"))
    assert np.array_equal(N_(rm.packbits(T(c["grid16x2"]), 1.5)), g["bits16x2"])
    # in-place variant and an unaligned / odd-sized view (scalar tail path)
    buf = torch.zeros(4096 + 3, dtype=torch.uint8, device=DEV)
    out = rm.packbits(T(c["grid32"]), c["thresh"], buf[3:])
    assert out.data_ptr() == buf[3:].data_ptr() and np.array_equal(N_(out), g["bits32"])
    odd = T(c["grid32"][:, :8 * 1001])
    assert np.array_equal(N_(rm.packbits(odd, c["thresh"])), g["bits32"][:1001])
    d = N_(rm.morton3D_dilation(T(c["grid32"])))
    assert np.array_equal(d, g["dil32"]) and np.array_equal(d, oracle.morton3D_dilation(c["grid32"]))
    assert np.array_equal(N_(rm.morton3D_dilation(T(c["grid16x2"]))), g["dil16x2"])
    s = N_(rm.sph_from_ray(T(c["sph_o"]), T(c["sph_d"]), c["radius"]))
    assert maxabs(s, g["sph"]) <= 1e-5 and maxabs(s, oracle.sph_from_ray(c["sph_o"], c["sph_d"], c["radius"])) <= 1e-5


def test_packbits_full_size_idempotent_roundtrip():
    """128^3 grid (BASELINE size): bitfield equals numpy.packbits and re-packing the unpacked bits is idempotent."""
    import raymarching as rm
    g = torch.rand(1, 128 ** 3, device=DEV) * 20
    bits = rm.packbits(g, 10.0)
    ref = np.packbits(N_(g).reshape(-1) > np.float32(10.0), bitorder="little")
    assert np.array_equal(N_(bits), ref)
    unpacked = torch.from_numpy(np.unpackbits(ref, bitorder="little").astype(np.float32)).to(DEV).view(1, -1)
    assert np.array_equal(N_(rm.packbits(unpacked, 0.5)), ref)
    dil = rm.morton3D_dilation(g)
    assert (dil >= g).all() and np.array_equal(N_(rm.morton3D_dilation(torch.zeros_like(g))), np.zeros((1, 128 ** 3), np.float32))


# ------------------------------------------------------------------------------------------------ marchers
@pytest.mark.parametrize("name", list(gc.MARCH_CASES))
def test_march_and_composite_train(oracle, name):
    import raymarching as rm
    c, g = gc.march_case(name), golden("march_" + name)
    ro, rd, bf = T(c["rays_o"]), T(c["rays_d"]), T(c["bitfield"])
    N, M = c["N"], c["M"]
    nears, fars = rm.near_far_from_aabb(ro, rd, T(c["aabb"]), c["min_near"])
    assert np.array_equal(N_(nears), g["nears"]) and np.array_equal(N_(fars), g["fars"])
    o_n, o_f = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    assert np.array_equal(o_n, g["nears"]) and np.array_equal(o_f, g["fars"])

    from radnerf_b200 import abi as L
    xyzs, dirs, deltas = torch.zeros(M, 3, device=DEV), torch.zeros(M, 3, device=DEV), torch.zeros(M, 2, device=DEV)
    rays = torch.empty(N, 3, dtype=torch.int32, device=DEV)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    t_noise = T(c["noises"])
    L.check(L.lib().rn_march_rays_train(L.ptr(ro), L.ptr(rd), L.ptr(bf), c["bound"], c["dt_gamma"], c["max_steps"], N, c["C"],
                                        c["H"], M, L.ptr(nears), L.ptr(fars), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas),
                                        L.ptr(rays), L.ptr(counter), L.ptr(t_noise), L.cur_stream()))
    ids, counts, kept, cx, cd, cdl = gc.canonical_rays(N_(rays), N_(xyzs), N_(dirs), N_(deltas), M)
    # bit-exact against the reference kernel: ids, per-ray sample counts, every emitted float, the counters
    assert np.array_equal(ids, np.arange(N)) and np.array_equal(counts, g["train_counts"]) and np.array_equal(kept, g["train_kept"])
    assert np.array_equal(cx, g["train_xyzs"]) and np.array_equal(cd, g["train_dirs"]) and np.array_equal(cdl, g["train_deltas"])
    assert np.array_equal(N_(counter), g["train_counter"]) and counter[0].item() == counts.sum() and counter[1].item() == N
    # offsets form an exact packing of [0, total)
    r = N_(rays)
    r = r[r[:, 2] > 0]
    r = r[np.argsort(r[:, 1])]
    assert r[0, 1] == 0 and np.array_equal(r[1:, 1], (r[:-1, 1] + r[:-1, 2]))
    # ... and against the CPU oracle
    ox, od, odl, orays, ocnt = oracle.march_rays_train(c["rays_o"], c["rays_d"], c["bound"], c["bitfield"], c["C"], c["H"], o_n,
                                                      o_f, c["noises"], M, c["dt_gamma"], c["max_steps"])
    _, ocounts, _, ocx, ocd, ocdl = gc.canonical_rays(orays, ox, od, odl, M)
    assert np.array_equal(ocounts, counts) and np.array_equal(ocx, cx) and np.array_equal(ocdl, cdl) and np.array_equal(ocd, cd)

    # composite on the canonical list
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    crays = np.stack([ids, offs, counts], 1).astype(np.int32)
    Mc = int(counts.sum())
    sig = T(c["sigmas"][:Mc]).requires_grad_(True)
    rgb = T(c["rgbs"][:Mc]).requires_grad_(True)
    amb = T(c["ambient"][:Mc]).requires_grad_(True)
    ws, ams, dp, im = rm.composite_rays_train(sig, rgb, amb, T(cdl), T(crays))
    for got, key in ((ws, "ct_ws"), (ams, "ct_amb"), (dp, "ct_depth"), (im, "ct_image")):
        assert maxabs(N_(got), g[key]) <= 1e-5 * max(1.0, float(np.abs(g[key]).max())), key
    o = oracle.composite_rays_train_forward(c["sigmas"][:Mc], c["rgbs"][:Mc], c["ambient"][:Mc], cdl, crays)
    for got, want in zip((ws, ams, dp, im), o):
        assert maxabs(N_(got), want) <= 1e-5 * max(1.0, float(np.abs(want).max()))
    torch.autograd.backward([ws, ams, im], [T(c["g_ws"]), T(c["g_amb"]), T(c["g_img"])])
    for got, key in ((sig.grad, "ct_gs"), (rgb.grad, "ct_gr"), (amb.grad, "ct_ga")):
        assert maxabs(N_(got), g[key]) <= 1e-5 * max(1.0, float(np.abs(g[key]).max())), key
    ogs, ogr, oga = oracle.composite_rays_train_backward(c["g_ws"], c["g_amb"], c["g_img"], c["sigmas"][:Mc], c["rgbs"][:Mc], cdl,
                                                         crays, o[0], o[3])
    assert maxabs(N_(sig.grad), ogs) <= 2e-5 * max(1.0, float(np.abs(ogs).max()))
    assert maxabs(N_(rgb.grad), ogr) <= 1e-5 and maxabs(N_(amb.grad), oga) == 0
    # march backward (camera optimisation path), identity ray order so slot == ray id
    gxyz = np.random.default_rng(5).standard_normal((Mc, 3)).astype(np.float32)
    gdir = np.random.default_rng(6).standard_normal((Mc, 3)).astype(np.float32)
    go, gd_ = torch.zeros(N, 3, device=DEV), torch.zeros(N, 3, device=DEV)
    t_gxyz, t_gdir, t_crays, t_cdl = T(gxyz), T(gdir), T(crays), T(cdl)  # keep the device buffers alive across the call
    L.check(L.lib().rn_march_rays_train_backward(L.ptr(t_gxyz), L.ptr(t_gdir), L.ptr(t_crays), L.ptr(t_cdl), N, Mc,
                                                 L.ptr(go), L.ptr(gd_), L.cur_stream()))
    assert maxabs(N_(go), g["mt_go"]) <= 1e-5 * max(1.0, float(np.abs(g["mt_go"]).max()))
    assert maxabs(N_(gd_), g["mt_gd"]) <= 1e-5 * max(1.0, float(np.abs(g["mt_gd"]).max()))


@pytest.mark.parametrize("name", list(gc.MARCH_CASES))
def test_march_and_composite_inference(oracle, name):
    import raymarching as rm
    c, g = gc.march_case(name), golden("march_" + name)
    ro, rd, bf = T(c["rays_o"]), T(c["rays_d"]), T(c["bitfield"])
    nears, fars = T(g["nears"]), T(g["fars"])
    na, ns = c["n_alive"], c["n_step"]
    from radnerf_b200 import abi as L
    Mi = na * ns
    Mi += 128 - (Mi % 128)
    alive, rays_t = T(c["rays_alive"]), nears.clone()
    xyzs, dirs, deltas = torch.zeros(Mi, 3, device=DEV), torch.zeros(Mi, 3, device=DEV), torch.zeros(Mi, 2, device=DEV)
    t_noise = T(c["infer_noises"])
    L.check(L.lib().rn_march_rays(na, ns, L.ptr(alive), L.ptr(rays_t), L.ptr(ro), L.ptr(rd), c["bound"], c["dt_gamma"],
                                  c["max_steps"], c["C"], c["H"], L.ptr(bf), L.ptr(nears), L.ptr(fars), L.ptr(xyzs),
                                  L.ptr(dirs), L.ptr(deltas), L.ptr(t_noise), L.cur_stream()))
    assert np.array_equal(N_(xyzs), g["inf_xyzs"]) and np.array_equal(N_(dirs), g["inf_dirs"]) and np.array_equal(N_(deltas), g["inf_deltas"])
    ox, od, odl = oracle.march_rays(na, ns, c["rays_alive"], g["nears"], c["rays_o"], c["rays_d"], c["bound"], c["bitfield"],
                                    c["C"], c["H"], g["nears"], g["fars"], 128, c["infer_noises"], c["dt_gamma"], c["max_steps"])
    assert np.array_equal(ox, g["inf_xyzs"]) and np.array_equal(od, g["inf_dirs"]) and np.array_equal(odl, g["inf_deltas"])
    # the wrapper allocates/pads exactly like the reference (zero noise here)
    wx, wd, wdl = rm.march_rays(na, ns, alive, rays_t, ro, rd, c["bound"], bf, c["C"], c["H"], nears, fars, 128, False,
                                c["dt_gamma"], c["max_steps"])
    assert wx.shape == (Mi, 3) and wdl.shape == (Mi, 2)

    ws, dp, im = T(c["ws0"]), T(c["depth0"]), T(c["image0"])
    rm.composite_rays(na, ns, alive, rays_t, T(c["sigmas"][:Mi]), T(c["rgbs"][:Mi]), deltas, ws, dp, im, 1e-2)
    assert np.array_equal(N_(alive), g["inf_alive"]) and np.array_equal(N_(rays_t), g["inf_rays_t"])
    for got, key in ((ws, "inf_ws"), (dp, "inf_depth"), (im, "inf_image")):
        assert maxabs(N_(got), g[key]) <= 1e-5 * max(1.0, float(np.abs(g[key]).max())), key
    o_alive, o_t = c["rays_alive"].copy(), g["nears"].copy()
    o_ws, o_dp, o_im = c["ws0"].copy(), c["depth0"].copy(), c["image0"].copy()
    oracle.composite_rays(na, ns, o_alive, o_t, c["sigmas"][:Mi], c["rgbs"][:Mi], odl, o_ws, o_dp, o_im, 1e-2)
    # alive flags may differ from the CPU only where T sits within an ulp of the threshold (different exp2 rounding)
    assert (o_alive != N_(alive)).mean() <= 1e-3
    assert maxabs(N_(im), o_im) <= 1e-5 * max(1.0, float(np.abs(o_im).max())) and maxabs(N_(ws), o_ws) <= 1e-5


def test_march_train_wrapper_allocation_rules():
    """raymarching.march_rays_train: M = N*max_steps when mean_count <= 0 (then trimmed to counter rounded UP by a full
    `align` even when aligned), M = mean_count rounded likewise otherwise; over-budget rays are dropped, never truncated."""
    import raymarching as rm
    c = gc.march_case("m_head")
    ro, rd, bf = T(c["rays_o"]), T(c["rays_d"]), T(c["bitfield"])
    nears, fars = rm.near_far_from_aabb(ro, rd, T(c["aabb"]), c["min_near"])
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, c["bound"], bf, c["C"], c["H"], nears, fars, counter, -1, False, 128,
                                                   False, c["dt_gamma"], c["max_steps"])
    total = counter[0].item()
    assert xyzs.shape[0] == total + 128 - total % 128 and rays.shape == (c["N"], 3) and counter[1].item() == c["N"]
    assert total == int(rays[:, 2].sum().item())
    # a budget that is too small: rays whose range would overflow are dropped as a whole
    budget = (total // 2) // 128 * 128
    counter.zero_()
    x2, d2, dl2, rays2 = rm.march_rays_train(ro, rd, c["bound"], bf, c["C"], c["H"], nears, fars, counter, budget - 128, False,
                                             128, False, c["dt_gamma"], c["max_steps"])
    assert x2.shape[0] == budget and counter[0].item() == total
    r2 = N_(rays2)
    over = (r2[:, 1].astype(np.int64) + r2[:, 2]) > budget
    assert over.any()
    dl2n = N_(dl2)
    for _, off, cnt in r2[(~over) & (r2[:, 2] > 0)][:200]:
        assert (dl2n[off:off + cnt, 0] > 0).all()
    ws, amb, dp, im = rm.composite_rays_train(torch.rand(budget, device=DEV), torch.rand(budget, 3, device=DEV),
                                              torch.rand(budget, device=DEV), dl2, rays2)
    dropped_ids = r2[over & (r2[:, 2] > 0)][:, 0]
    assert not N_(im)[dropped_ids].any() and not N_(ws)[dropped_ids].any()


def test_empty_inputs_are_noops():
    import raymarching as rm
    e3 = torch.zeros(0, 3, device=DEV)
    n, f = rm.near_far_from_aabb(e3, e3, torch.tensor([-1., -1, -1, 1, 1, 1], device=DEV), 0.05)
    assert n.shape == (0,) and f.shape == (0,)
    assert rm.morton3D(torch.zeros(0, 3, dtype=torch.int32, device=DEV)).shape == (0,)
    from gridencoder import GridEncoder
    enc = GridEncoder(input_dim=2, num_levels=4, log2_hashmap_size=10).to(DEV)
    assert enc(torch.zeros(0, 2, device=DEV)).shape == (0, 8)


# ------------------------------------------------------------------------------------------------ freq / SH
def test_freq_and_sh_encoders(oracle):
    from freqencoder import FreqEncoder
    from shencoder import SHEncoder
    c, g = gc.enc_case(), golden("enc")
    for key, D, deg in (("freq2", 2, 10), ("freq6", 6, 4)):
        enc = FreqEncoder(input_dim=D, degree=deg)
        assert enc.output_dim == D + 2 * D * deg
        x = T(c[key]).requires_grad_(True)
        y = enc(x)
        # same intrinsic (__sinf on scalbnf(x,f)+phase) as the reference -> agreement far inside 1e-5
        assert maxabs(N_(y), g[key + "_out"]) <= 1e-5
        y.backward(T(c["g" + key]))
        assert maxabs(N_(x.grad), g[key + "_gin"]) <= 1e-5 * max(1.0, float(np.abs(g[key + "_gin"]).max()))
        # CPU oracle uses libm sinf: sin.approx loses absolute accuracy as |arg| grows (arguments reach 2^9 * 0.8 rad)
        o = oracle.freq_encode_forward(c[key], deg)
        assert maxabs(N_(y), o) <= 2e-3
        ogi = oracle.freq_encode_backward(c["g" + key], N_(y), D, deg)
        assert maxabs(N_(x.grad), ogi) <= 1e-4 * max(1.0, float(np.abs(ogi).max()))
    for deg, gkey in ((4, "gsh4"), (8, "gsh8")):
        enc = SHEncoder(degree=deg)
        d = T(c["dirs"]).requires_grad_(True)
        y = enc(d)
        ref = g[f"sh{deg}_out"]
        assert maxabs(N_(y), ref) <= 1e-5 * max(1.0, float(np.abs(ref).max()))
        o, ody = oracle.sh_encode_forward(c["dirs"], deg, True)
        assert maxabs(N_(y), o) <= 1e-5 * max(1.0, float(np.abs(o).max()))
        assert maxabs(ref, o) <= 1e-5 * max(1.0, float(np.abs(o).max()))  # oracle pinned to the reference kernel
        assert maxabs(g[f"sh{deg}_dy"], ody) <= 1e-5 * max(1.0, float(np.abs(ody).max()))
        y.backward(T(c[gkey]))
        gin = g[f"sh{deg}_gin"]
        assert maxabs(N_(d.grad), gin) <= 2e-5 * max(1.0, float(np.abs(gin).max()))
    assert SHEncoder(degree=4)(T(c["dirs"])).shape == (256, 16)
