"""Shared inputs / deterministic weights of the network parity case: used by
tests/golden/make_network_golden.py (REFERENCE NeRFNetwork on the CPU) and tests/test_network_parity.py (ours)."""
import numpy as np
import torch


def fill_parameters(module, table_scale=0.5):
    """overwrite every parameter with values that only depend on its NAME and shape (both classes have identical state-dict keys)"""
    with torch.no_grad():
        for i, (name, p) in enumerate(sorted(module.named_parameters(), key=lambda kv: kv[0])):
            g = torch.Generator().manual_seed(1000 + i)
            if name.endswith("embeddings"):
                v = (torch.rand(p.shape, generator=g) * 2 - 1) * table_scale
            elif p.dim() >= 2:
                fan_in = int(np.prod(p.shape[1:]))
                v = (torch.rand(p.shape, generator=g) * 2 - 1) * (1.5 / np.sqrt(fan_in))
            else:
                v = (torch.rand(p.shape, generator=g) * 2 - 1) * 0.1
            p.copy_(v.to(p.device, p.dtype))


def inputs(n=1536, n_torso=1024):
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * torch.tensor([0.9, 0.45, 0.9])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    auds = torch.randn(8, 44, 16, generator=g) * 3.0
    eye = torch.tensor([[0.25]])
    xy = torch.rand(n_torso, 2, generator=g) * 2 - 1
    poses = torch.tensor([[0.05, -0.12, 0.02, 0.07, 3.38, -0.23]])
    return dict(x=x, d=d, auds=auds, eye=eye, xy=xy, poses=poses)
