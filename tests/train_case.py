"""Inputs of the whole-training-step parity case (BASELINE configs[3] at a size the CPU finishes in seconds), shared by
tests/golden/make_train_golden.py (the REFERENCE's NeRFNetwork + NeRFRenderer.run_cuda + Trainer.train_step on the CPU) and
tests/test_train_parity.py (our mirror on the CPU operators and on the CUDA operators)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rad-nerf_b200"))
HW, N_RAYS, FRAME = 64, 1024, 3
GLOBAL_STEP, ITERS, LAMBDA_AMB = 300, 1000, 0.1          # nerf/utils.py:803: lambda = min(step / iters, 1) * opt.lambda_amb
TOP_ROWS = 2048                                           # table-gradient rows kept in the fixture (largest |gradient|)


def lambda_amb():
    return min(GLOBAL_STEP / ITERS, 1.0) * LAMBDA_AMB


TABLES = ("encoder.embeddings", "encoder_ambient.embeddings", "torso_encoder.embeddings")


def install_occupancy(net, torso=False):
    from radnerf_b200 import synthetic as syn
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    with torch.no_grad():
        net.density_grid.copy_(torch.from_numpy(grid).to(net.density_grid.device))
        net.mean_density = float(np.clip(grid, 0, None).mean())
        net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(net.mean_density, net.density_thresh))).to(net.density_bitfield.device))
        if torso:
            tg = syn.torso_density_grid(128)
            net.density_grid_torso.copy_(torch.from_numpy(tg).to(net.density_grid_torso.device))
            net.mean_density_torso = float(tg.mean())


def install_head_occupancy(net):
    install_occupancy(net, torso=False)


def batch():
    """the keys NeRFDataset.collate hands Trainer.train_step (nerf/provider.py:250-290), CPU tensors"""
    from radnerf_b200 import synthetic as syn
    b = syn.batch_to(syn.training_batch(HW, HW, N_RAYS, frame_index=FRAME), "cpu")
    b["images"] = b["rgb"]
    b["bg_torso_color"] = b["rgb"]        # the torso phase's target (nerf/utils.py:730): same synthetic colours
    return b


def noise():
    """start offsets of the marcher for perturb=True: fixed, so that every implementation marches the same samples"""
    return np.random.default_rng(77).random(N_RAYS, dtype=np.float32)


def summarise(named_grads, tables=TABLES):
    """what the fixture keeps of a set of gradients: small tensors whole; a table as its level-free summary -- the TOP_ROWS rows
    with the largest gradient (indices + values), the column sums and the L2 norm"""
    out = {}
    for name, g in named_grads.items():
        g = np.asarray(g, np.float32)
        if name in tables:
            rows = np.argsort(-np.abs(g).max(axis=1), kind="stable")[:TOP_ROWS]
            out[name + "/rows"] = rows.astype(np.int32)
            out[name + "/values"] = g[rows]
            out[name + "/colsum"] = g.astype(np.float64).sum(0)
            out[name + "/norm"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
            out[name + "/nonzero_rows"] = np.int64((np.abs(g).max(axis=1) > 0).sum())
        else:
            out[name] = g
    return out
