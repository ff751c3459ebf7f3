"""`trunc_exp` -- the density activation of nerf/network.py (reference: activation.py): y = exp(x) evaluated in fp32 even under
autocast; the gradient uses exp(clamp(x, -15, 15)) so that a huge pre-activation cannot produce an infinite gradient.
The fused renderer applies the same function inside its sigma epilogue."""
import torch
from torch.amp import custom_bwd, custom_fwd


class TruncExpFn(torch.autograd.Function):
    GRAD_CLAMP = 15.0

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, pre):
        ctx.save_for_backward(pre)
        return pre.exp()

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        (pre,) = ctx.saved_tensors
        return dy * pre.clamp(-TruncExpFn.GRAD_CLAMP, TruncExpFn.GRAD_CLAMP).exp()


trunc_exp = TruncExpFn.apply
