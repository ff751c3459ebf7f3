"""FreqEncoder -- sin/cos positional encoding on libradnerf_b200.  Drop-in for /root/reference/freqencoder/freq.py
(same class/attributes, `_freq_encoder` autograd contract: inputs cast to fp32, outputs fp32)."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi as _L


class _freq_encoder(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)  # force float32 for better precision
    def forward(ctx, inputs, degree, output_dim):
        # inputs: [B, input_dim] float -> [B, output_dim] float
        if not inputs.is_cuda:
            inputs = inputs.cuda()
        inputs = inputs.contiguous()
        B, input_dim = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        _L.check(_L.lib().rn_freq_encode_forward(_L.ptr(inputs), B, input_dim, degree, output_dim, _L.ptr(outputs),
                                                 _L.cur_stream()))
        ctx.save_for_backward(inputs, outputs)
        ctx.dims = [B, input_dim, degree, output_dim]
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        grad = grad.contiguous()
        inputs, outputs = ctx.saved_tensors
        B, input_dim, degree, output_dim = ctx.dims
        grad_inputs = torch.empty_like(inputs)
        _L.check(_L.lib().rn_freq_encode_backward(_L.ptr(grad), _L.ptr(outputs), B, input_dim, degree, output_dim,
                                                  _L.ptr(grad_inputs), _L.cur_stream()))
        return grad_inputs, None, None


freq_encode = _freq_encoder.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        # inputs: [..., input_dim] -> [..., output_dim]
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = freq_encode(inputs, self.degree, self.output_dim)
        outputs = outputs.reshape(prefix_shape + [self.output_dim])
        return outputs
