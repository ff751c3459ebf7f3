"""GridEncoder -- multi-resolution hash / tiled grid encoding on libradnerf_b200 (sm_100a).

Drop-in for /root/reference/gridencoder/grid.py: same class, constructor arguments, attributes, state-dict keys
(`embeddings` fp32 [rows, C], `offsets` int32 [L+1]) and autograd/AMP contract (`_grid_encode`, grid.py:24-89).
Differences are internal: the CUDA kernel writes [B, L*C] rows directly (no [L,B,C] buffer + permute copy,
grid.py:47,57), consumes the incoming gradient in that layout (no permute copy, grid.py:75) and accumulates the
table gradient in fp32 (the reference accumulates fp16 atomics under autocast, grid.py:77).
"""
import numpy as np

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi as _L

_gridtype_to_id = {
    'hash': 0,
    'tiled': 1,
}

_interp_to_id = {
    'linear': 0,
    'smoothstep': 1,
}

_RN_F32, _RN_F16 = 0, 1
_LAYOUT_LBC, _LAYOUT_BLC = 0, 1


def _dtype_id(t):
    if t.dtype == torch.float32:
        return _RN_F32
    if t.dtype == torch.float16:
        return _RN_F16
    raise RuntimeError(f"GridEncoder: embeddings must be float32 or float16, got {t.dtype}")


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0):
        # inputs: [B, D], float in [0, 1]; embeddings: [sO, C]; offsets: [L + 1] int32.  RETURN: [B, L * C]
        _L.require_cuda(inputs, embeddings, offsets)
        if inputs.dtype != torch.float32:
            raise RuntimeError("GridEncoder: inputs must be float32 (coordinates are never reduced in precision)")
        inputs = inputs.contiguous()

        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = np.log2(per_level_scale)
        H = base_resolution

        emb_dtype_in = embeddings.dtype
        # manual autocast handling, as the reference (grid.py:41-44): half tables only when C is even
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        embeddings = embeddings.contiguous()
        offsets = offsets.contiguous()

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None

        _L.check(_L.lib().rn_grid_encode_forward(
            _L.ptr(inputs), _L.ptr(embeddings), _L.ptr(offsets), _L.ptr(outputs), B, D, C, L, float(S), H,
            _L.ptr(dy_dx), gridtype, int(bool(align_corners)), interpolation, _dtype_id(embeddings), _LAYOUT_BLC,
            _L.cur_stream()))

        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype, interpolation]
        ctx.align_corners = align_corners
        ctx.emb_dtype_in = emb_dtype_in
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation = ctx.dims
        align_corners = ctx.align_corners

        grad = grad.contiguous()  # [B, L*C], consumed in place of the reference's [L,B,C] permuted copy
        if grad.dtype != embeddings.dtype:
            grad = grad.to(embeddings.dtype)

        # fp32 accumulation target for the scatter-add, whatever the table dtype
        grad_embeddings = torch.zeros(embeddings.shape, device=embeddings.device, dtype=torch.float32)
        grad_inputs = torch.empty_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None

        _L.check(_L.lib().rn_grid_encode_backward(
            _L.ptr(grad), _L.ptr(inputs), _L.ptr(embeddings), _L.ptr(offsets), _L.ptr(grad_embeddings), B, D, C, L,
            float(S), H, _L.ptr(dy_dx), _L.ptr(grad_inputs), gridtype, int(bool(align_corners)), interpolation,
            _dtype_id(embeddings), _LAYOUT_BLC, _RN_F32, _L.cur_stream()))

        if dy_dx is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        if grad_embeddings.dtype != ctx.emb_dtype_in:
            grad_embeddings = grad_embeddings.to(ctx.emb_dtype_in)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear'):
        super().__init__()

        # the finest resolution desired at the last level, if provided, overrides per_level_scale
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        # level table: rows per level capped at 2^log2_hashmap_size and rounded up to a multiple of 8 (grid.py:118-127)
        offsets = []
        offset = 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        offsets = torch.from_numpy(np.array(offsets, dtype=np.int32))
        self.register_buffer('offsets', offsets)

        self.n_params = offsets[-1] * level_dim

        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))

        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} interpolation={self.interpolation}"

    def forward(self, inputs, bound=1):
        # inputs: [..., input_dim], normalized real world positions in [-bound, bound]
        # return: [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)  # map to [0, 1]

        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)

        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        outputs = outputs.view(prefix_shape + [self.output_dim])
        return outputs

    # always run in float precision!
    @torch.autocast(device_type="cuda", enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        # inputs: [..., input_dim], float in [-b, b], location to calculate TV loss.
        D = self.input_dim
        C = self.embeddings.shape[1]
        L = self.offsets.shape[0] - 1
        S = np.log2(self.per_level_scale)
        H = self.base_resolution

        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = (inputs + bound) / (2 * bound)
            inputs = inputs.view(-1, self.input_dim)
            B = inputs.shape[0]

        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')

        inputs = inputs.to(self.embeddings.dtype).contiguous()
        _L.check(_L.lib().rn_grad_total_variation(
            _L.ptr(inputs), _L.ptr(self.embeddings), _L.ptr(self.embeddings.grad), _L.ptr(self.offsets), float(weight),
            B, D, C, L, float(S), H, self.gridtype_id, int(bool(self.align_corners)), _dtype_id(self.embeddings),
            _L.cur_stream()))
