"""SURVEY 8(a) rows a16 / a17 -- the network mirror (radnerf_b200.model.NeRFNetwork: AudioNet, AudioAttNet, ambient / sigma /
colour / torso MLPs over the three grid encoders) against golden vectors produced by the REFERENCE's own NeRFNetwork class
(tests/golden/make_network_golden.py imported nerf/network.py from /root/reference and ran it on the CPU with the same
name-derived weights).  CPU: our model on the oracle-backed operators, fp32.  GPU: the CUDA operators in fp32 and under the
fp16 autocast the reference runs with."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden", "network.npz")


def _outputs(device, ops=None, autocast=False):
    from network_case import fill_parameters, inputs
    from radnerf_b200.model import NeRFNetwork, Options
    net = NeRFNetwork(Options(torso=True, smooth_lips=False, fp16=autocast), ops=ops).eval().to(device)
    fill_parameters(net)
    c = {k: v.to(device) for k, v in inputs().items()}
    out = {}
    with torch.no_grad(), torch.autocast(device if device != "cpu" else "cpu", dtype=torch.float16, enabled=autocast):
        enc_a = net.encode_audio(c["auds"])
        out["enc_a"] = enc_a
        sigma, color, ambient = net(c["x"], c["d"], enc_a, net.individual_codes[0], c["eye"])
        out.update(sigma=sigma, color=color, ambient=ambient)
        den = net.density(c["x"], enc_a, c["eye"])
        out.update(density_sigma=den["sigma"], density_geo=den["geo_feat"])
        alpha, rgb, deform = net.forward_torso(c["xy"], c["poses"], enc_a, net.individual_codes_torso[0])
        out.update(torso_alpha=alpha, torso_color=rgb, torso_deform=deform)
    return {k: v.float().cpu().numpy() for k, v in out.items()}


def _compare(got, tol):
    g = np.load(GOLD)
    assert set(got) == set(g.files)
    for k in g.files:
        a, b = got[k].reshape(g[k].shape), g[k]
        err = float(np.abs(a - b).max())
        assert err <= tol * max(1.0, float(np.abs(b).max())), (k, err)


def test_network_mirror_cpu_port_matches_the_reference_class():
    from oracle.cpu_backend import CPUOps
    _compare(_outputs("cpu", CPUOps()), 1e-6)     # same torch-CPU layers, same oracle encoders: only op-order noise


@pytest.mark.gpu
def test_network_mirror_cuda_fp32_matches_the_reference_class():
    _compare(_outputs("cuda"), 1e-5)              # north_star: <= 1e-5 for fp32


@pytest.mark.gpu
def test_network_mirror_cuda_fp16_autocast_matches_the_reference_class():
    _compare(_outputs("cuda", autocast=True), 2e-3)   # fp16 tables and layers (the reference's -O mode) vs its fp32 evaluation
