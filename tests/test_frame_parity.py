"""SURVEY 8(a) row a18 -- a whole head+torso frame against a golden frame rendered by the REFERENCE's own
NeRFNetwork.render / NeRFRenderer.run_cuda (tests/golden/make_frame_golden.py ran nerf/renderer.py:158-316 from
/root/reference on the CPU over the oracle's operators, fp32).  Ours: the CPU port (same operators), the CUDA op-by-op path
in fp32, and the fused sm_100a frame (fp16 tables / layers, the reference's -O mode)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden", "frame.npz")


def _render(device, ops=None, fp16=False, path="ops"):
    from network_case import fill_parameters
    from frame_case import frame_inputs, install_occupancy
    from radnerf_b200.model import NeRFNetwork, Options
    opt = Options(torso=True, smooth_lips=False, fp16=fp16)
    net = NeRFNetwork(opt, ops=ops).eval().to(device)
    fill_parameters(net)
    if ops is not None:
        scales = np.load(os.path.join(ROOT, "tests", "golden", "grid_g3_f32.npz"))["scales"]
        for e in (net.encoder, net.encoder_ambient, net.torso_encoder):
            e.device_scales = scales
    install_occupancy(net)
    f = {k: v.to(device) for k, v in frame_inputs().items()}
    with torch.no_grad(), torch.autocast("cuda" if device != "cpu" else "cpu", dtype=torch.float16, enabled=fp16):
        out = net.render(f["rays_o"], f["rays_d"], f["auds"], f["bg_coords"], f["poses"], eye=f["eye"], index=0, bg_color=None,
                         perturb=False, path=path, **opt.render_kwargs())
    return {k: out[k].float().cpu().numpy().reshape(-1, 3) if k in ("image", "torso_color") else out[k].float().cpu().numpy().reshape(-1)
            for k in ("image", "depth", "torso_alpha", "torso_color")}


def _compare(got, tol):
    g = np.load(GOLD)
    assert int((np.abs(g["image"] - 1).max(-1) > 1e-3).sum()) == 1188     # the golden frame shows head and torso
    for k in ("image", "depth", "torso_alpha", "torso_color"):
        err = float(np.abs(got[k] - g[k]).max())
        assert err <= tol, (k, err)


def test_frame_cpu_port_matches_the_reference_renderer():
    from oracle.cpu_backend import CPUOps
    _compare(_render("cpu", CPUOps()), 1e-5)


@pytest.mark.gpu
def test_frame_cuda_ops_fp32_matches_the_reference_renderer():
    saved = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # see tests/test_network_parity.py
    try:
        _compare(_render("cuda"), 5e-4)     # torso: __sinf frequency encoder vs libm in the golden run (see test_network_parity)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved


@pytest.mark.gpu
def test_frame_fused_fp16_matches_the_reference_renderer():
    # the fused sm_100a frame in the reference's fp16 mode against the reference's fp32 frame: white-noise tables of
    # amplitude 0.5 make this the worst case for fp16 (see test_network_parity); colours are in [0, 1]
    _compare(_render("cuda", fp16=True, path="fused"), 2e-2)
