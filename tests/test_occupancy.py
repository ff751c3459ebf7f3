"""SURVEY 8(a) row a19 -- occupancy maintenance (mark_untrained_grid, update_extra_state) against golden vectors produced
by the REFERENCE's own NeRFRenderer code (tests/golden/make_occupancy_golden.py ran nerf/renderer.py:318-501 from
/root/reference on the CPU over the oracle's operators).  The CPU test runs our model on the oracle-backed CPU operators,
the GPU test on the CUDA library."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden", "occupancy.npz")


def _run(device, ops=None):
    from occupancy_case import analytic_sigma, case_inputs
    from radnerf_b200.model import NeRFNetwork, Options
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=False, ind_dim=0, att=0), ops=ops).to(device)
    m.density = lambda x, enc_a, e=None: {"sigma": analytic_sigma(x)}
    m.encode_audio = lambda a: None
    m.aud_features = torch.zeros(4, 1, 16)
    m.eye_area = torch.full((4, 1), 0.25)
    c = case_inputs()
    m.mark_untrained_grid(c["poses"], c["intrinsics"])
    untrained = (m.density_grid < 0).cpu().numpy().reshape(-1)
    orig = torch.rand_like
    torch.rand_like = lambda t, **kw: torch.full_like(t, 0.5)   # no in-cell jitter: the result is a function of the inputs
    try:
        m.local_step = 3
        m.step_counter[:3, 0] = torch.tensor([100, 200, 301], dtype=torch.int32)
        m.update_extra_state()
        grid1 = m.density_grid.clone().cpu()
        md1 = m.mean_density
        m.update_extra_state()
    finally:
        torch.rand_like = orig
    return m, untrained, grid1, md1


def _check(m, untrained, grid1, md1, exact):
    g = np.load(GOLD)
    want_untrained = np.unpackbits(g["untrained_bits"], bitorder="little").astype(bool)
    n_diff = int((untrained != want_untrained).sum())
    # CPU vs the reference's CPU run: same arithmetic, must be identical.  GPU: the 3x3 products of the frustum test are
    # summed in another order, cells exactly on a frustum plane may flip (measured: a handful of 2 M).
    assert n_diff == 0 if exact else n_diff <= 64, n_diff
    assert int(want_untrained.sum()) == 1046808
    both = ~(untrained | want_untrained)
    a, b = grid1.numpy().reshape(-1)[both], g["grid_after_1"].astype(np.float32).reshape(-1)[both]
    assert np.abs(a - b).max() <= 2e-3 * max(1.0, np.abs(b).max())     # golden grid is stored in fp16
    assert abs(md1 - float(g["mean_density_1"])) <= 1e-4 * float(g["mean_density_1"])
    assert abs(m.mean_density - float(g["mean_density_2"])) <= 1e-4 * float(g["mean_density_2"])
    s = m.density_grid.cpu().numpy().reshape(-1)[::4099]
    keep = (s >= 0) == (g["grid_after_2_sample"] >= 0)
    assert keep.mean() > 0.999 and np.abs(s[keep] - g["grid_after_2_sample"][keep]).max() <= 1e-4
    bits = np.unpackbits(m.density_bitfield.cpu().numpy() ^ g["bitfield_2"]).sum()
    assert bits == 0 if exact else bits <= 64, bits      # cells within rounding of the threshold
    assert m.mean_count == int(g["mean_count"]) and m.local_step == 0 and m.iter_density == 2


def test_occupancy_maintenance_cpu_port_matches_the_reference_code():
    from oracle.cpu_backend import CPUOps
    _check(*_run("cpu", CPUOps()), exact=True)


@pytest.mark.gpu
def test_occupancy_maintenance_cuda_matches_the_reference_code():
    _check(*_run("cuda"), exact=False)


def test_torso_occupancy_update_cpu_port_matches_the_reference_code():
    """the TORSO branch of update_extra_state (nerf/renderer.py:455-501) with the real torso network: golden from the reference's
    own NeRFNetwork(torso=True) on the CPU (tests/golden/make_torso_occupancy_golden.py); our mirror queries all 128^2 cells in
    one vectorised pass instead of the reference's chunk loop -- same grid (<= 1e-6), same mean density, same mean_count, and the
    head's density grid is left alone while the torso trains"""
    import random
    from network_case import fill_parameters
    from occupancy_case import torso_case
    from oracle.cpu_backend import CPUOps
    from radnerf_b200.model import NeRFNetwork, Options
    g = np.load(os.path.join(ROOT, "tests", "golden", "occupancy_torso.npz"))
    m = NeRFNetwork(Options(torso=True, smooth_lips=False, fp16=False, exp_eye=True), ops=CPUOps())
    fill_parameters(m)
    scales = np.load(os.path.join(ROOT, "tests", "golden", "grid_g3_f32.npz"))["scales"]
    for e in (m.encoder, m.encoder_ambient, m.torso_encoder):
        e.device_scales = scales
    c = torso_case()
    m.aud_features, m.eye_area, m.poses = c["aud_features"], c["eye_area"], c["poses"]
    m.density_grid_torso.copy_(c["grid0"])
    head_before = m.density_grid.clone()
    m.local_step = 2
    m.step_counter[:2, 0] = torch.tensor([700, 901], dtype=torch.int32)
    orig = torch.rand_like
    torch.rand_like = lambda t, **kw: torch.full_like(t, 0.5)
    random.seed(c["seed"])
    try:
        m.update_extra_state()
        g1, md1, mc = m.density_grid_torso.clone(), m.mean_density_torso, m.mean_count
        m.update_extra_state()
    finally:
        torch.rand_like = orig
    assert np.abs(g1.numpy() - g["grid_after_1"]).max() <= 1e-6
    assert np.abs(m.density_grid_torso.numpy() - g["grid_after_2"]).max() <= 1e-6
    assert abs(md1 - float(g["mean_density_1"])) <= 1e-6 and abs(m.mean_density_torso - float(g["mean_density_2"])) <= 1e-6
    assert mc == int(g["mean_count"]) == 800 and m.local_step == 0
    assert torch.equal(m.density_grid, head_before)
