"""Multi-resolution hash / tiled grid encoder on libradnerf_b200 (sm_100a).

Public surface of the reference's gridencoder/grid.py: `GridEncoder(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2,
base_resolution=16, log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
interpolation='linear')` with the state-dict keys `embeddings` (fp32 [rows, C]) and `offsets` (int32 [L+1]), `forward(inputs,
bound=1)`, `grad_total_variation(...)`, and the functional `grid_encode(...)` with the reference's positional arguments.

What differs is below the surface: the kernel writes [B, L*C] rows directly (the reference fills [L, B, C] and permute-copies,
grid.py:47,57), the backward consumes the gradient in that same layout and accumulates the table gradient in fp32 (the
reference uses fp16 atomics under autocast, grid.py:77)."""
import math

import numpy as np
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi

GRID_TYPES = {'hash': 0, 'tiled': 1}
INTERPOLATIONS = {'linear': 0, 'smoothstep': 1}
_gridtype_to_id, _interp_to_id = GRID_TYPES, INTERPOLATIONS   # names other code may import
F32, F16 = 0, 1            # RN_F32 / RN_F16
ROWS_BLC = 1               # RN_LAYOUT_BLC: [B, L*C]


def _table_dtype_id(t):
    try:
        return {torch.float32: F32, torch.float16: F16}[t.dtype]
    except KeyError:
        raise RuntimeError(f"GridEncoder: embeddings must be float32 or float16, got {t.dtype}") from None


def level_offsets(input_dim, num_levels, level_dim, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """first row of every level (+ total): a level has min(2^log2_hashmap_size, points^D) rows rounded up to a multiple of 8,
    points = ceil(H * s^l) (+1 unless align_corners) -- the table geometry every RAD-NeRF checkpoint is laid out in"""
    cap, rows, starts = 2 ** log2_hashmap_size, 0, []
    for lvl in range(num_levels):
        points = int(np.ceil(base_resolution * per_level_scale ** lvl)) + (0 if align_corners else 1)
        starts.append(rows)
        rows += 8 * math.ceil(min(cap, points ** input_dim) / 8)
    return np.asarray(starts + [rows], dtype=np.int32)


class GridEncodeFn(torch.autograd.Function):
    """(x in [0,1]^D, table, offsets) -> features [B, L*C]; optional d(features)/dx for input gradients"""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, table, offsets, per_level_scale, base_resolution, want_input_grad=False, gridtype=0, align_corners=False,
                interpolation=0):
        abi.require_cuda(x, table, offsets)
        x = x.float()   # coordinates are evaluated in fp32 whatever arrives (the reference's grid.py:35 `inputs.float()` under autocast)
        param_dtype = table.dtype
        if torch.is_autocast_enabled() and table.shape[1] % 2 == 0:   # the reference's own autocast rule (grid.py:41-44)
            table = table.half()
        x, table, offsets = x.contiguous(), table.contiguous(), offsets.contiguous()
        n, dim = x.shape
        levels, feats = offsets.numel() - 1, table.shape[1]
        log2_scale = float(np.log2(per_level_scale))
        y = table.new_empty(n, levels * feats)
        jac = table.new_empty(n, levels * dim * feats) if want_input_grad else None
        abi.call("rn_grid_encode_forward", x, table, offsets, y, n, dim, feats, levels, log2_scale, base_resolution, jac, gridtype,
                 int(bool(align_corners)), interpolation, _table_dtype_id(table), ROWS_BLC)
        ctx.save_for_backward(x, table, offsets, jac)
        ctx.geometry = (log2_scale, base_resolution, gridtype, int(bool(align_corners)), interpolation, param_dtype)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, table, offsets, jac = ctx.saved_tensors
        log2_scale, base_resolution, gridtype, align, interpolation, param_dtype = ctx.geometry
        n, dim = x.shape
        levels, feats = offsets.numel() - 1, table.shape[1]
        dy = dy.contiguous().to(table.dtype)
        dtable = torch.zeros(table.shape, device=table.device, dtype=torch.float32)     # fp32 scatter target
        dx = torch.empty_like(x, dtype=table.dtype) if jac is not None else None
        abi.call("rn_grid_encode_backward", dy, x, table, offsets, dtable, n, dim, feats, levels, log2_scale, base_resolution, jac, dx,
                 gridtype, align, interpolation, _table_dtype_id(table), ROWS_BLC, F32)
        return (None if dx is None else dx.to(x.dtype)), dtable.to(param_dtype), None, None, None, None, None, None, None


grid_encode = GridEncodeFn.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False, interpolation='linear'):
        super().__init__()
        if desired_resolution is not None:   # finest resolution given: derive the geometric growth factor from it
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution, self.log2_hashmap_size = per_level_scale, base_resolution, log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, GRID_TYPES[gridtype]
        self.interpolation, self.interp_id = interpolation, INTERPOLATIONS[interpolation]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size
        offsets = torch.from_numpy(level_offsets(input_dim, num_levels, level_dim, per_level_scale, base_resolution,
                                                 log2_hashmap_size, align_corners))
        self.register_buffer('offsets', offsets)
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.uniform_(self.embeddings, -1e-4, 1e-4)

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} "
                f"interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        unit = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)     # [-bound, bound] -> [0, 1]
        feats = grid_encode(unit, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, unit.requires_grad,
                            self.gridtype_id, self.align_corners, self.interp_id)
        return feats.view(*inputs.shape[:-1], self.output_dim)

    @torch.autocast(device_type="cuda", enabled=False)   # total variation is always evaluated in fp32
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """adds weight * d(TV)/d(table) at `inputs` (or B random points) INTO embeddings.grad -- call between backward() and step()"""
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        if inputs is None:
            pts = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            pts = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
        pts = pts.to(self.embeddings.dtype).contiguous()
        abi.call("rn_grad_total_variation", pts, self.embeddings, self.embeddings.grad, self.offsets, float(weight), pts.shape[0],
                 self.input_dim, self.embeddings.shape[1], self.offsets.numel() - 1, float(np.log2(self.per_level_scale)),
                 self.base_resolution, self.gridtype_id, int(bool(self.align_corners)), _table_dtype_id(self.embeddings))
