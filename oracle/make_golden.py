"""Produce the golden vectors under tests/golden/ by running the REFERENCE's own compiled CUDA kernels (oracle/_ref/*.so,
built from /root/reference by oracle/build_ref.py) on the seeded cases of tests/golden_cases.py.

Needs a GPU:   gpurun -- python oracle/make_golden.py gpurun_out/golden      (then copy *.npz into tests/golden/)
Only the compiled extension modules are used -- none of the reference's Python -- so this runs on the GPU box where
/root/reference does not exist.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases as gc  # noqa: E402


def load_ref():
    import _gridencoder, _raymarching_face, _freqencoder, _shencoder  # noqa: E401
    return _gridencoder, _raymarching_face, _freqencoder, _shencoder


dev = torch.device("cuda")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
N_ = lambda t: t.detach().cpu().numpy()


def run_grid(G, name):
    c = gc.grid_case(name)
    x, table, offs = T(c["inputs"]), T(c["table"]), T(c["offsets"])
    B, D, C, L, H = c["B"], c["D"], c["C"], c["L"], c["H"]
    S = float(np.log2(c["per_level_scale"]))
    out = torch.empty(L, B, C, device=dev, dtype=table.dtype)
    dy = torch.empty(B, L * D * C, device=dev, dtype=table.dtype) if c["dy"] else None
    G.grid_encode_forward(x, table, offs, out, B, D, C, L, S, H, dy, c["gridtype"], c["align"], c["interp"])
    res = {"out": N_(out.permute(1, 0, 2).reshape(B, L * C).contiguous())}
    if dy is not None:
        res["dy_dx"] = N_(dy)
    grad = T(c["grad"]).view(B, L, C).permute(1, 0, 2).contiguous()
    ge = torch.zeros_like(table)
    gi = torch.zeros(B, D, device=dev, dtype=table.dtype) if c["dy"] else None
    G.grid_encode_backward(grad, x, table, offs, ge, B, D, C, L, S, H, dy, gi, c["gridtype"], c["align"], c["interp"])
    ge = N_(ge.float())
    res["grad_emb_rows"] = ge[c["bwd_rows"]]
    res["grad_emb_level_sums"] = np.stack([ge[c["offsets"][l]:c["offsets"][l + 1]].astype(np.float64).sum(0) for l in range(L)])
    if gi is not None:
        res["grad_inputs"] = N_(gi.float())
    # level geometry as the device computes it
    lv = torch.arange(L, device=dev, dtype=torch.float32)
    res["scales"] = N_(torch.exp2(lv * np.float32(S)) * np.float32(H) - 1.0)
    if not c["half"] and c["dy"]:
        # total-variation gradient (fp32 tables), inputs = the same points
        g = torch.zeros_like(table)
        G.grad_total_variation(x, table, g, offs, 1e-3, B, D, C, L, S, H, c["gridtype"], c["align"])
        g = N_(g)
        res["tv_rows"] = g[c["bwd_rows"]]
    return res


def run_utils(R):
    c = gc.util_case()
    res = {}
    coords = T(c["coords"]); ind = torch.empty(coords.shape[0], dtype=torch.int32, device=dev)
    R.morton3D(coords, coords.shape[0], ind); res["morton"] = N_(ind)
    indices = T(c["indices"]); co = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=dev)
    R.morton3D_invert(indices, indices.shape[0], co); res["invert"] = N_(co)
    g = T(c["grid32"]); bf = torch.empty(g.numel() // 8, dtype=torch.uint8, device=dev)
    R.packbits(g, g.numel() // 8, c["thresh"], bf); res["bits32"] = N_(bf)
    gd = torch.empty_like(g); R.morton3D_dilation(g, 1, 32, gd); res["dil32"] = N_(gd)
    g2 = T(c["grid16x2"]); gd2 = torch.empty_like(g2); R.morton3D_dilation(g2, 2, 16, gd2); res["dil16x2"] = N_(gd2)
    bf2 = torch.empty(g2.numel() // 8, dtype=torch.uint8, device=dev); R.packbits(g2, g2.numel() // 8, 1.5, bf2); res["bits16x2"] = N_(bf2)
    o, d = T(c["sph_o"]), T(c["sph_d"]); sc = torch.empty(o.shape[0], 2, device=dev)
    R.sph_from_ray(o, d, c["radius"], o.shape[0], sc); res["sph"] = N_(sc)
    return res


def run_march(R, name):
    c = gc.march_case(name)
    res = {}
    ro, rd, aabb, bf = T(c["rays_o"]), T(c["rays_d"]), T(c["aabb"]), T(c["bitfield"])
    N, M = c["N"], c["M"]
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    R.near_far_from_aabb(ro, rd, aabb, N, c["min_near"], nears, fars)
    res["nears"], res["fars"] = N_(nears), N_(fars)
    # ---- training march
    xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev); counter = torch.zeros(2, dtype=torch.int32, device=dev)
    R.march_rays_train(ro, rd, bf, c["bound"], c["dt_gamma"], c["max_steps"], N, c["C"], c["H"], M, nears, fars, xyzs, dirs,
                       deltas, rays, counter, T(c["noises"]))
    ids, counts, kept, cx, cd, cdl = gc.canonical_rays(N_(rays), N_(xyzs), N_(dirs), N_(deltas), M)
    res.update(train_ids=ids, train_counts=counts, train_kept=kept, train_xyzs=cx, train_dirs=cd, train_deltas=cdl,
               train_counter=N_(counter))
    # ---- training composite on the canonical (ray-id ordered) sample list
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    crays = np.stack([ids, offs, counts], 1).astype(np.int32)
    Mc = int(counts.sum())
    sig, rgb, amb = T(c["sigmas"][:Mc]), T(c["rgbs"][:Mc]), T(c["ambient"][:Mc])
    dl, cr = T(cdl), T(crays)
    ws, ams, dp, im = (torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev),
                       torch.empty(N, 3, device=dev))
    R.composite_rays_train_forward(sig, rgb, amb, dl, cr, Mc, N, 1e-4, ws, ams, dp, im)
    res.update(ct_ws=N_(ws), ct_amb=N_(ams), ct_depth=N_(dp), ct_image=N_(im))
    gs, gr, ga = torch.zeros(Mc, device=dev), torch.zeros(Mc, 3, device=dev), torch.zeros(Mc, device=dev)
    R.composite_rays_train_backward(T(c["g_ws"]), T(c["g_amb"]), T(c["g_img"]), sig, rgb, amb, dl, cr, ws, ams, im, Mc, N,
                                    1e-4, gs, gr, ga)
    res.update(ct_gs=N_(gs), ct_gr=N_(gr), ct_ga=N_(ga))
    gxyz = torch.from_numpy(np.random.default_rng(5).standard_normal((Mc, 3)).astype(np.float32)).to(dev)
    gdir = torch.from_numpy(np.random.default_rng(6).standard_normal((Mc, 3)).astype(np.float32)).to(dev)
    go, gd_ = torch.zeros(N, 3, device=dev), torch.zeros(N, 3, device=dev)
    R.march_rays_train_backward(gxyz, gdir, cr, dl, N, Mc, go, gd_)
    res.update(mt_go=N_(go), mt_gd=N_(gd_))
    # ---- inference march + composite (slot layout is deterministic)
    na, ns = c["n_alive"], c["n_step"]
    Mi = na * ns; Mi += 128 - (Mi % 128)
    alive = T(c["rays_alive"]); rays_t = nears.clone()
    ix, idr, idl = torch.zeros(Mi, 3, device=dev), torch.zeros(Mi, 3, device=dev), torch.zeros(Mi, 2, device=dev)
    R.march_rays(na, ns, alive, rays_t, ro, rd, c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"], bf, nears, fars,
                 ix, idr, idl, T(c["infer_noises"]))
    res.update(inf_xyzs=N_(ix), inf_dirs=N_(idr), inf_deltas=N_(idl))
    ws0, d0, im0 = T(c["ws0"]), T(c["depth0"]), T(c["image0"])
    R.composite_rays(na, ns, 1e-2, alive, rays_t, T(c["sigmas"][:Mi]), T(c["rgbs"][:Mi]), idl, ws0, d0, im0)
    res.update(inf_alive=N_(alive), inf_rays_t=N_(rays_t), inf_ws=N_(ws0), inf_depth=N_(d0), inf_image=N_(im0))
    return res


def run_enc(F, Sx):
    c = gc.enc_case()
    res = {}
    for key, D, deg in (("freq2", 2, 10), ("freq6", 6, 4)):
        x = T(c[key]); B = x.shape[0]; Cc = D + 2 * D * deg
        out = torch.empty(B, Cc, device=dev); F.freq_encode_forward(x, B, D, deg, Cc, out)
        gi = torch.zeros(B, D, device=dev); F.freq_encode_backward(T(c["g" + key]), out, B, D, deg, Cc, gi)
        res[key + "_out"], res[key + "_gin"] = N_(out), N_(gi)
    d = T(c["dirs"]); B = d.shape[0]
    for deg, gkey in ((4, "gsh4"), (8, "gsh8")):
        out = torch.empty(B, deg * deg, device=dev); dy = torch.empty(B, 3 * deg * deg, device=dev)
        Sx.sh_encode_forward(d, out, B, 3, deg, dy)
        gi = torch.zeros(B, 3, device=dev); Sx.sh_encode_backward(T(c[gkey]), d, B, 3, deg, dy, gi)
        res[f"sh{deg}_out"], res[f"sh{deg}_dy"], res[f"sh{deg}_gin"] = N_(out), N_(dy), N_(gi)
    return res


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    G, R, F, Sx = load_ref()
    for name in gc.GRID_CASES:
        np.savez_compressed(os.path.join(outdir, f"grid_{name}.npz"), **run_grid(G, name))
        print("grid", name, "ok")
    np.savez_compressed(os.path.join(outdir, "utils.npz"), **run_utils(R)); print("utils ok")
    for name in gc.MARCH_CASES:
        np.savez_compressed(os.path.join(outdir, f"march_{name}.npz"), **run_march(R, name)); print("march", name, "ok")
    np.savez_compressed(os.path.join(outdir, "enc.npz"), **run_enc(F, Sx)); print("enc ok")
    torch.cuda.synchronize()
    with open(os.path.join(outdir, "PROVENANCE.txt"), "w") as f:
        f.write("generated by oracle/make_golden.py from the reference's compiled CUDA extensions (oracle/_ref) on %s, torch %s\n"
                % (torch.cuda.get_device_name(0), torch.__version__))
    print("done ->", outdir)


if __name__ == "__main__":
    main()
