"""CPU tests of the oracle itself and of the host-side logic (no GPU needed)."""
import numpy as np
import pytest

import golden_cases as gc


def test_structural_constants_of_the_reference(oracle):
    # level offsets / table sizes recorded in the reference (gridencoder/grid.py:127,134; nerf/utils.py:1490-1491)
    offs3, pls = oracle.grid_offsets(3, 16, 2, 16, 16, 2048)
    assert offs3.tolist() == [0, 4920, 18744, 51512, 117048, 182584, 248120, 313656, 379192, 444728, 510264, 575800,
                              641336, 706872, 772408, 837944, 903480]
    assert abs(pls - 1.381912879967776) < 1e-15 and abs(np.log2(pls) - 0.4666666666666666) < 1e-15
    offs2, _ = oracle.grid_offsets(2, 16, 2, 16, 16, 2048)
    assert offs2[-1] == 555520
    sc, res = oracle.grid_level_geometry(np.log2(pls), 16, 16)
    assert res.tolist() == [16, 23, 31, 43, 59, 81, 112, 154, 213, 295, 407, 562, 777, 1073, 1483, 2048]
    assert sc[0] == 15.0 and sc[-1] == 2047.0


def test_morton_identities(oracle):
    c = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [127, 127, 127], [5, 9, 77]], np.int32)
    m = oracle.morton3D(c)
    assert m[:4].tolist() == [1, 2, 4, 2097151]
    assert np.array_equal(oracle.morton3D_invert(m), c)
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (5000, 3)).astype(np.int32)
    assert np.array_equal(oracle.morton3D_invert(oracle.morton3D(c)), c)
    # matches the vectorised generator used for synthetic inputs
    from radnerf_b200 import synthetic as syn
    assert np.array_equal(oracle.morton3D(c).astype(np.uint32), syn.morton3D_np(c[:, 0], c[:, 1], c[:, 2]))


def test_packbits_and_dilation(oracle):
    rng = np.random.default_rng(1)
    g = rng.random((1, 16 ** 3), dtype=np.float32)
    assert np.array_equal(oracle.packbits(g, 0.5), np.packbits(g.reshape(-1) > 0.5, bitorder="little"))
    d = oracle.morton3D_dilation(g)
    # brute force in xyz order
    idx = oracle.morton3D(np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3).astype(np.int32))
    vol = np.zeros((16, 16, 16), np.float32)
    vol.reshape(-1)[:] = g[0, idx]
    pad = np.pad(vol, 1, constant_values=-np.inf)
    ref = np.maximum.reduce([pad[1:-1, 1:-1, 1:-1], pad[2:, 1:-1, 1:-1], pad[:-2, 1:-1, 1:-1], pad[1:-1, 2:, 1:-1],
                             pad[1:-1, :-2, 1:-1], pad[1:-1, 1:-1, 2:], pad[1:-1, 1:-1, :-2]])
    assert np.array_equal(d[0, idx], ref.reshape(-1))


def test_near_far(oracle):
    o = np.array([[0, 0, -3], [0, 0, -3], [5, 5, 5]], np.float32)
    d = np.array([[0, 0, 1], [0, 1, 0], [0, 0, 1]], np.float32)
    n, f = oracle.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.05)
    assert n[0] == 2 and f[0] == 4
    assert n[2] == np.finfo(np.float32).max and f[2] == n[2]  # a miss is FLT_MAX on both


def test_grid_encode_against_numpy_restatement(oracle):
    """2-D dense levels: bilinear interpolation written directly with numpy."""
    c = gc.grid_case("g2_f32_dy")
    out, dy = oracle.grid_encode_forward(c["inputs"], c["table"], c["offsets"], c["per_level_scale"], c["H"], True, 1, False, 0)
    sc, res = oracle.grid_level_geometry(np.float32(np.log2(c["per_level_scale"])), c["H"], c["L"])
    x = c["inputs"].astype(np.float64)
    for l in range(9):  # dense 2-D levels
        pos = x * float(sc[l]) + 0.5
        p0 = np.floor(pos)
        f = pos - p0
        p0 = p0.astype(np.int64)
        acc = 0
        for dx in (0, 1):
            for dyy in (0, 1):
                w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dyy else 1 - f[:, 1])
                idx = (p0[:, 0] + dx) + (p0[:, 1] + dyy) * (int(res[l]) + 1) + int(c["offsets"][l])
                acc = acc + w[:, None] * c["table"][np.clip(idx, 0, len(c["table"]) - 1)].astype(np.float64)
        ok = (c["inputs"] >= 0).all(1) & (c["inputs"] <= 1).all(1)
        assert np.abs(out[ok, 2 * l:2 * l + 2] - acc[ok]).max() < 2e-5
    assert not out[:2].any() and not dy[:2].any()  # out-of-range rows


def test_grid_backward_is_the_adjoint_of_forward(oracle):
    """<grad, forward(table)> == <backward(grad), table> since the encoding is linear in the table."""
    for name in ("g3_hash_sm_f32", "g2_align_c1", "g4_f32"):
        c = gc.grid_case(name)
        out, _ = oracle.grid_encode_forward(c["inputs"], c["table"], c["offsets"], c["per_level_scale"], c["H"], False,
                                            c["gridtype"], c["align"], c["interp"])
        ge, _ = oracle.grid_encode_backward(c["grad"], c["inputs"], c["offsets"], c["per_level_scale"], c["H"],
                                            c["table"].shape[0], c["C"], None, c["gridtype"], c["align"], c["interp"])
        lhs = float((out.astype(np.float64) * c["grad"].astype(np.float64)).sum())
        rhs = float((ge * c["table"].astype(np.float64)).sum())
        assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


def test_half_mode_rounds_per_corner(oracle):
    c = gc.grid_case("g3_f16")
    out16, _ = oracle.grid_encode_forward(c["inputs"], c["table"], c["offsets"], c["per_level_scale"], c["H"], False, 1, False, 0)
    out32, _ = oracle.grid_encode_forward(c["inputs"], c["table"].astype(np.float32), c["offsets"], c["per_level_scale"], c["H"], False, 1, False, 0)
    assert out16.dtype == np.float16
    d = np.abs(out16.astype(np.float32) - out32).max()
    assert 0 < d < 3e-3  # differs from fp32 accumulation by a few half-ulps, never more


def test_march_train_properties(oracle):
    for name in gc.MARCH_CASES:
        c = gc.march_case(name)
        n, f = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
        x, d, dl, rays, cnt = oracle.march_rays_train(c["rays_o"], c["rays_d"], c["bound"], c["bitfield"], c["C"], c["H"], n, f,
                                                     c["noises"], c["M"], c["dt_gamma"], c["max_steps"])
        assert cnt[1] == c["N"] and cnt[0] == rays[:, 2].sum() and rays[:, 2].max() <= c["max_steps"]
        assert (rays[:, 2] > 0).any()
        tot = cnt[0]
        assert (np.abs(x[:tot]) <= c["bound"]).all() and (dl[:tot, 0] > 0).all()
        # t is increasing along every ray and stays below far; dirs are the ray direction
        for rid, off, k in rays[rays[:, 2] > 1][:50]:
            t = dl[off:off + k, 1]
            assert (np.diff(t) > 0).all() and t[-1] - dl[off + k - 1, 0] < f[rid]
            assert np.array_equal(d[off], c["rays_d"][rid])
        # the inference marcher, run in chunks of 3 from rays_t, reproduces exactly the training marcher's samples
        rid = int(rays[np.argmax(rays[:, 2]), 0])
        off, k = rays[rays[:, 0] == rid][0, 1:]
        rt = n.copy()
        rt[rid] = x_t0 = np.float32(n[rid])  # zero noise below, so the start is `near`
        alive, nz, got = np.array([rid], np.int32), np.zeros(1, np.float32), []
        x0, _, dl0, rays0, _ = oracle.march_rays_train(c["rays_o"], c["rays_d"], c["bound"], c["bitfield"], c["C"], c["H"], n, f,
                                                       np.zeros(c["N"], np.float32), c["M"], c["dt_gamma"], c["max_steps"])
        off0, k0 = rays0[rays0[:, 0] == rid][0, 1:]
        while sum(len(g) for g in got) < k0:
            xs, ds, dls = oracle.march_rays(1, 3, alive, rt, c["rays_o"], c["rays_d"], c["bound"], c["bitfield"], c["C"], c["H"], n, f,
                                            -1, nz, c["dt_gamma"], c["max_steps"])
            kk = int((dls[:, 0] > 0).sum())
            got.append(xs[:kk])
            if kk < 3:
                break
            rt[rid] = dls[kk - 1, 1]
        got = np.concatenate(got)[:k0]
        assert len(got) == k0 and np.array_equal(got, x0[off0:off0 + k0])


def test_composite_backward_matches_finite_differences(oracle):
    rng = np.random.default_rng(3)
    N, K = 6, 7
    rays = np.stack([np.arange(N), np.arange(N) * K, np.full(N, K)], 1).astype(np.int32)
    M = N * K
    sig = (rng.random(M) * 3).astype(np.float32)
    rgb = rng.random((M, 3)).astype(np.float32)
    amb = rng.random(M).astype(np.float32)
    dl = np.stack([np.full(M, 0.05, np.float32), np.cumsum(np.full(M, 0.05, np.float32))], 1)
    g_ws, g_amb, g_img = rng.standard_normal(N).astype(np.float32), rng.standard_normal(N).astype(np.float32), rng.standard_normal((N, 3)).astype(np.float32)
    ws, am, dp, im = oracle.composite_rays_train_forward(sig, rgb, amb, dl, rays)
    gs, gr, ga = oracle.composite_rays_train_backward(g_ws, g_amb, g_img, sig, rgb, dl, rays, ws, im)

    def loss(s, c):
        w, a, _, i = oracle.composite_rays_train_forward(s, c, amb, dl, rays)
        return float((w.astype(np.float64) * g_ws).sum() + (i.astype(np.float64) * g_img).sum() + (a.astype(np.float64) * g_amb).sum())
    eps = 1e-2
    for k in (0, 5, 20, M - 1):
        sp, sm = sig.copy(), sig.copy()
        sp[k] += eps; sm[k] -= eps
        fd = (loss(sp, rgb) - loss(sm, rgb)) / (2 * eps)
        assert abs(fd - gs[k]) < 5e-3 * max(1.0, abs(fd))
        cp, cm = rgb.copy(), rgb.copy()
        cp[k, 1] += eps; cm[k, 1] -= eps
        fd = (loss(sig, cp) - loss(sig, cm)) / (2 * eps)
        assert abs(fd - gr[k, 1]) < 5e-3 * max(1.0, abs(fd))
    assert np.array_equal(ga, np.repeat(g_amb, K))
    # empty / dropped rays give zeros
    rays2 = rays.copy(); rays2[1, 2] = 0; rays2[2, 1] = M  # empty ray, out-of-budget ray
    w2, a2, d2, i2 = oracle.composite_rays_train_forward(sig, rgb, amb, dl, rays2)
    assert w2[1] == 0 and not i2[2].any()


def test_composite_rays_inference_semantics(oracle):
    """T = 1 - weights_sum, termination on a zero delta or T < T_thresh marks the ray dead (raymarching.cu:979-1022)."""
    N, ns = 4, 3
    alive = np.arange(N, dtype=np.int32)
    rays_t = np.zeros(N, np.float32)
    dl = np.zeros((N * ns, 2), np.float32)
    dl[:, 0] = 0.1; dl[:, 1] = np.tile([1, 2, 3], N)
    dl[1 * ns + 1:2 * ns] = 0  # ray 1 runs out after one sample
    sig = np.full(N * ns, 1.0, np.float32); sig[2 * ns] = 1e4  # ray 2 becomes opaque at its first sample
    rgb = np.ones((N * ns, 3), np.float32)
    ws, dp, im = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    oracle.composite_rays(N, ns, alive, rays_t, sig, rgb, dl, ws, dp, im, 1e-2)
    assert alive.tolist() == [0, -1, -1, 3] and rays_t[0] == 3 and rays_t[1] == 0
    a = 1 - np.exp(-0.1)
    assert abs(ws[0] - (1 - (1 - a) ** 3)) < 1e-6 and abs(ws[1] - a) < 1e-6 and abs(ws[2] - 1.0) < 1e-4
    assert np.allclose(im[0], ws[0], atol=1e-6)


def test_freq_and_sh_oracle_basics(oracle):
    x = np.array([[0.25, -0.5]], np.float32)
    y = oracle.freq_encode_forward(x, 3)[0]
    exp = np.concatenate([x[0]] + [f(x[0] * 2 ** k) for k in range(3) for f in (np.sin, np.cos)])
    assert np.abs(y - exp).max() < 1e-6
    g = np.random.default_rng(0).standard_normal((1, 14)).astype(np.float32)
    gi = oracle.freq_encode_backward(g, y[None], 2, 3)[0]
    eps = 1e-3
    for d in range(2):
        xp, xm = x.copy(), x.copy(); xp[0, d] += eps; xm[0, d] -= eps
        fd = ((oracle.freq_encode_forward(xp, 3) - oracle.freq_encode_forward(xm, 3)) * g).sum() / (2 * eps)
        assert abs(fd - gi[d]) < 1e-2
    d = np.array([[0.3, 0.5, 0.8], [0, 0, 1]], np.float32)
    o, dy = oracle.sh_encode_forward(d, 4, True)
    # first bands in closed form: 1/(2 sqrt(pi)), -sqrt(3/4pi) y, sqrt(3/4pi) z, -sqrt(3/4pi) x
    k = np.sqrt(3 / (4 * np.pi))
    assert np.allclose(o[0, :4], [0.28209479, -k * 0.5, k * 0.8, -k * 0.3], atol=1e-6)
    assert abs(o[1, 6] - np.sqrt(5 / np.pi) / 4 * 2) < 1e-6  # (3z^2-1) at z = 1
    # orthonormality on the sphere (Monte-Carlo) up to degree 8
    v = np.random.default_rng(1).standard_normal((200000, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    Y, _ = oracle.sh_encode_forward(v.astype(np.float32), 8)
    G = Y.astype(np.float64).T @ Y.astype(np.float64) * (4 * np.pi / len(v))
    assert np.abs(G - np.eye(64)).max() < 0.05
    # gradient by finite differences
    eps = 1e-3
    for ax in range(3):
        dp, dm = d.copy(), d.copy(); dp[:, ax] += eps; dm[:, ax] -= eps
        fd = (oracle.sh_encode_forward(dp, 8)[0] - oracle.sh_encode_forward(dm, 8)[0]) / (2 * eps)
        an = oracle.sh_encode_forward(d, 8, True)[1].reshape(2, 3, 64)[:, ax]
        assert np.abs(fd - an).max() < 2e-2


def test_synthetic_scene_generators():
    from radnerf_b200 import synthetic as syn
    pose = syn.orbit_pose(10.0)
    R = pose[:3, :3]
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-6) and abs(np.linalg.norm(pose[:3, 3]) - 3.35) < 1e-5
    ro, rd = syn.get_rays(pose, syn.intrinsics_for(64, 64), 64, 64)
    assert ro.shape == (4096, 3) and np.allclose(np.linalg.norm(rd, axis=1), 1, atol=1e-6)
    centre = rd[32 * 64 + 32]
    assert np.dot(centre, -pose[:3, 3] / 3.35) > 0.999  # the central ray looks at the origin
    bank = syn.audio_feature_bank(20, 44, 16)
    assert syn.audio_window(bank, 0).shape == (8, 44, 16) and not syn.audio_window(bank, 0)[:4].any()
    assert syn.audio_window(bank, 19).shape == (8, 44, 16) and not syn.audio_window(bank, 19)[5:].any()
    assert syn.get_bg_coords(8, 8).shape == (64, 2)
    g = syn.head_density_grid(32)
    assert 0.005 < (g > 0).mean() < 0.1
