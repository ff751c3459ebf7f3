"""Frequency (sin / cos) positional encoding on libradnerf_b200.

Public surface of the reference's freqencoder/freq.py: `FreqEncoder(input_dim=3, degree=4)` with `.input_dim`, `.degree`,
`.output_dim = input_dim * (1 + 2 * degree)`, `forward(inputs, **kw)`, and the functional `freq_encode(x, degree, width)`.
Layout per row (freqencoder.cu:30-58): [x | per octave k: sin(2^k x), cos(2^k x)], computed in fp32 regardless of autocast.
The kernels use the fast `__sinf` path exactly as the reference's `-use_fast_math` build does."""
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi


class FreqEncodeFn(torch.autograd.Function):
    """forward: rn_freq_encode_forward; backward: rn_freq_encode_backward (needs the forward's outputs: d sin = cos)"""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, n_octaves, width):
        x, _ = abi.rows(x.cuda() if not x.is_cuda else x)
        y = x.new_empty(x.shape[0], width)
        abi.call("rn_freq_encode_forward", x, x.shape[0], x.shape[1], n_octaves, width, y)
        ctx.save_for_backward(x, y)
        ctx.meta = (n_octaves, width)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        n_octaves, width = ctx.meta
        dx = torch.empty_like(x)
        abi.call("rn_freq_encode_backward", dy.contiguous(), y, x.shape[0], x.shape[1], n_octaves, width, dx)
        return dx, None, None


freq_encode = FreqEncodeFn.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim * (1 + 2 * degree)

    def extra_repr(self):
        return f"input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def __repr__(self):
        return "FreqEncoder: " + self.extra_repr()

    def forward(self, inputs, **kwargs):
        flat = inputs.reshape(-1, self.input_dim)
        return freq_encode(flat, self.degree, self.output_dim).view(*inputs.shape[:-1], self.output_dim)
