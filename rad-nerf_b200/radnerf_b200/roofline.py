"""Live roofline measurement for bench.py: times the dominant kernel(s) alone with CUDA events on the launching stream
(L2 flushed between launches) and relates ALGORITHMIC bytes to the measured B200 HBM peak (MEASURED_PEAKS.json).

Algorithmic bytes per unit (restated in DESIGN.md):
  grid encode fwd : 4*D + L*2^D*C*s + L*C*s per sample   (3-D fp16: 588 B, 2-D fp16: 328 B)
  march_rays      : 44 B per alive ray + 32 B per emitted sample
  composite_rays  : 24 B per sample + 56 B per alive ray
"""
import json
import os

import torch

from . import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def peaks():
    p = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


_flush_buf = None


def flush_l2():
    global _flush_buf
    if _flush_buf is None:
        _flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush_buf.zero_()


def time_kernel(fn, iters=20, warmup=3, flush=True):
    """average device time of fn() in ms, each launch timed separately on the current stream"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def frame_samples(model, f, kw, min_samples=1 << 18):
    """occupied sample positions of a real frame (marching all rays for max_steps samples), tiled up to min_samples"""
    import raymarching as rm
    ro, rd = f["ro"][0], f["rd"][0]
    nears, fars = rm.near_far_from_aabb(ro, rd, model.aabb_infer, model.min_near)
    N = ro.shape[0]
    alive = torch.arange(N, dtype=torch.int32, device=ro.device)
    xyzs, dirs, deltas = rm.march_rays(N, kw["max_steps"], alive, nears.clone(), ro, rd, model.bound, model.density_bitfield,
                                       model.cascade, model.grid_size, nears, fars, -1, False, kw["dt_gamma"], kw["max_steps"])
    keep = deltas[:, 0] > 0
    x = xyzs[keep]
    while x.shape[0] < min_samples:
        x = torch.cat([x, x], 0)
    return x.contiguous(), dirs[keep], nears, fars


def grid_fwd_roofline(enc, x01, half=True):
    import numpy as np
    B, D = x01.shape
    L, Cc = enc.num_levels, enc.level_dim
    emb = enc.embeddings.detach().to(torch.half if half else torch.float32).contiguous()
    out = torch.empty(B, L * Cc, device=x01.device, dtype=emb.dtype)
    S = float(np.log2(enc.per_level_scale))

    def launch():
        abi.check(abi.lib().rn_grid_encode_forward(abi.ptr(x01), abi.ptr(emb), abi.ptr(enc.offsets), abi.ptr(out), B, D, Cc, L, S,
                                                   enc.base_resolution, None, enc.gridtype_id, int(enc.align_corners),
                                                   enc.interp_id, 1 if half else 0, 1, abi.cur_stream()))
    ms = time_kernel(launch)
    s = 2 if half else 4
    bytes_per = 4 * D + L * (2 ** D) * Cc * s + L * Cc * s
    return ms, bytes_per * B, B


def measure(model, f, bg_local, kw, path):
    hbm, tflops, src = peaks()
    x, dirs, nears, fars = frame_samples(model, f, kw)
    x01 = ((x + model.bound) / (2 * model.bound)).contiguous()
    kernels = []
    ms, nbytes, B = grid_fwd_roofline(model.encoder, x01, half=True)
    g3 = {"kernel": "grid_forward_kernel<half,3,2> (3-D hash-grid encode, fp16 table)", "bound": "hbm", "units": B,
          "unit_name": "samples", "bytes_per_unit": nbytes // B, "ms": ms, "achieved": nbytes / ms / 1e6, "peak": hbm,
          "unit": "GB/s", "frac": nbytes / ms / 1e6 / hbm, "gunits_per_s": B / ms / 1e6}
    kernels.append(g3)
    amb = (torch.rand(B, 2, device=x.device)).contiguous()
    ms2, nbytes2, _ = grid_fwd_roofline(model.encoder_ambient, amb, half=True)
    kernels.append({"kernel": "grid_forward_kernel<half,2,2> (2-D ambient/torso grid encode)", "bound": "hbm", "units": B,
                    "bytes_per_unit": nbytes2 // B, "ms": ms2, "achieved": nbytes2 / ms2 / 1e6, "peak": hbm, "unit": "GB/s",
                    "frac": nbytes2 / ms2 / 1e6 / hbm})
    extra = []
    try:
        from . import frame
        extra = frame.roofline_entries(model, f, bg_local, kw, hbm, tflops)
    except Exception:
        pass
    kernels += extra
    dom = extra[0] if (path == "fused" and extra) else g3
    roof = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"],
            "traffic": dom.get("traffic"), "traffic_note": dom.get("traffic_note"), "kernel": dom["kernel"], "peak_source": src + " (MEASURED_PEAKS.json)" if src == "measured" else src,
            "algorithmic_per_launch": dom.get("bytes_per_unit", 0) * dom.get("units", 0), "launch_ms": dom["ms"]}
    if dom.get("issue"):
        roof["issue"] = dom["issue"]
    return roof, kernels
