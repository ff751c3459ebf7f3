"""Golden vectors for one WHOLE training step (BASELINE configs[3]; SURVEY 8(a) rows a3, a9, a10, a16, a17 through autograd),
produced by the REFERENCE's own classes on the CPU of this container:

    NeRFNetwork (nerf/network.py) . train()  ->  NeRFRenderer.run_cuda, training branch (nerf/renderer.py:207-236)
    ->  Trainer.train_step (nerf/utils.py:718-808, called unbound on a bare object)  ->  loss.backward()

once for the head phase and once for the torso phase (`--torso`: the loss is taken on `torso_color` against `bg_torso_color`, only the
torso branch -- deformation MLP, 2-D torso grid, torso MLP, torso individual codes -- receives gradients; nerf/utils.py:730, 744, 787).

`raymarching` / `encoding` / `activation` are bound to the oracle's CPU operators with autograd (oracle/cpu_backend.py,
CPUOps(train=True); the kernels underneath are pinned to the reference's CUDA kernels by tests/test_oracle_golden.py), weights
come from tests/network_case.fill_parameters, the batch / marcher noise / lambda schedule from tests/train_case.py; fp32.

    python tests/golden/make_train_golden.py      ->  tests/golden/train_step.npz, train_step_torso.npz   (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from network_case import fill_parameters   # noqa: E402
import train_case as tc   # noqa: E402


def main():
    from oracle import cpu_backend
    ops = cpu_backend.CPUOps(train=True)
    rm = types.ModuleType("raymarching")
    for n in ("morton3D", "morton3D_dilation", "packbits", "near_far_from_aabb", "march_rays", "composite_rays", "march_rays_train",
              "composite_rays_train"):
        setattr(rm, n, getattr(ops.rm, n))
    sys.modules["raymarching"] = rm
    enc = types.ModuleType("encoding")
    enc.get_encoder = cpu_backend.get_encoder
    sys.modules["encoding"] = enc
    act = types.ModuleType("activation")
    act.trunc_exp = ops.trunc_exp
    sys.modules["activation"] = act
    for name in ("trimesh", "tensorboardX", "matplotlib", "matplotlib.pyplot", "mcubes", "imageio", "lpips"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "torch_ema" not in sys.modules:
        m = types.ModuleType("torch_ema")
        m.ExponentialMovingAverage = object
        sys.modules["torch_ema"] = m
    sys.path.append("/root/reference")
    from nerf.network import NeRFNetwork
    from nerf.utils import Trainer
    from radnerf_b200.model import Options

    scales = np.load(os.path.join(HERE, "grid_g3_f32.npz"))["scales"]
    for phase in ("head", "torso"):
        o = Options(torso=(phase == "torso"), smooth_lips=False, fp16=False, exp_eye=True)
        opt = types.SimpleNamespace(**{**vars(o), "test_train": False, "color_space": "srgb", "patch_size": 1, "finetune_lips": False,
                                       "iters": tc.ITERS, "lambda_amb": tc.LAMBDA_AMB})
        net = NeRFNetwork(opt).train()
        fill_parameters(net)
        for e in ([net.encoder, net.encoder_ambient] + ([net.torso_encoder] if phase == "torso" else [])):
            e.device_scales = scales
        tc.install_occupancy(net, torso=(phase == "torso"))
        cpu_backend.TRAIN_NOISE = tc.noise()
        me = types.SimpleNamespace(opt=opt, model=net, criterion=torch.nn.MSELoss(reduction="none"), global_step=tc.GLOBAL_STEP,
                                   flip_finetune_lips=False)
        b = tc.batch()
        pred, truth, loss = Trainer.train_step(me, b)
        loss.backward()
        counter = net.step_counter[0].numpy().copy()
        grads = {n: p.grad.numpy() for n, p in net.named_parameters() if p.grad is not None}
        res = tc.summarise(grads)
        res["loss"] = np.float64(loss.item())
        res["pred_rgb"] = pred.detach().numpy().reshape(-1, 3)
        res["counter"] = counter
        res["grad_names"] = np.array(sorted(grads))
        np.savez_compressed(os.path.join(HERE, "train_step.npz" if phase == "head" else "train_step_torso.npz"), **res)
        print(phase, "loss", float(loss.detach()), "counter", counter.tolist(), "tensors with a gradient:", len(grads))
        for n in tc.TABLES:
            if n + "/norm" in res:
                print("  ", n, "norm", float(res[n + "/norm"]), "nonzero rows", int(res[n + "/nonzero_rows"]))


if __name__ == "__main__":
    main()
