"""Real spherical-harmonics direction encoding on libradnerf_b200.

Public surface of the reference's shencoder/sphere_harmonics.py: `SHEncoder(input_dim=3, degree=4)` (degree 1..8,
`output_dim = degree^2`), `forward(inputs, size=1)` which scales directions by 1/size first, and the functional
`sh_encode(x, degree, need_input_grad)`.  fp32 regardless of autocast; the Jacobian is only materialised when the caller
asks for input gradients (shencoder.cu:27-438)."""
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi


class SHEncodeFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, dirs, degree, need_input_grad=False):
        dirs, _ = abi.rows(dirs)
        n, d = dirs.shape
        y = dirs.new_empty(n, degree * degree)
        jac = dirs.new_empty(n, d * degree * degree) if need_input_grad else None
        abi.call("rn_sh_encode_forward", dirs, y, n, d, degree, jac)
        ctx.save_for_backward(dirs, jac)
        ctx.degree = degree
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        dirs, jac = ctx.saved_tensors
        if jac is None:
            return None, None, None
        ddirs = torch.zeros_like(dirs)
        abi.call("rn_sh_encode_backward", dy.contiguous(), dirs, dirs.shape[0], dirs.shape[1], ctx.degree, jac, ddirs)
        return ddirs, None, None


sh_encode = SHEncodeFn.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        if input_dim != 3:
            raise AssertionError("SH encoder only support input dim == 3")
        if not 1 <= degree <= 8:
            raise AssertionError("SH encoder only supports degree in [1, 8]")
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree * degree

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        unit = (inputs / size).reshape(-1, self.input_dim)
        return sh_encode(unit, self.degree, unit.requires_grad).view(*inputs.shape[:-1], self.output_dim)
