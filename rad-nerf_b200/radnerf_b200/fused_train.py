"""Fused training step of the head network: host side of rn_head_train_* (include/radnerf_b200.h, csrc/head_train_*.cu).

`head_forward(model, xyzs, dirs, enc_a, ind_code, eye)` has the contract of NeRFNetwork.forward on a batch of march_rays_train
samples (nerf/network.py:222-283: `sigma [M], color [M,3], ambient [M,2]`) and is differentiable w.r.t. everything the reference's
is: the two hash tables, the eight Linear weights, the audio code `enc_a` (-> AudioNet / AudioAttNet through torch autograd) and
the individual code.  Forward = ONE kernel, backward = two tcgen05 kernels + the two table-scatter kernels + a 21 k-element
reduction; the reference runs 16 cuBLAS GEMMs and ~300 elementwise launches for the same work.

Per-frame-constant network inputs (audio code, eye, individual code) never become per-sample columns: their first-layer products are
hoisted into three 64-vectors before the forward (as the inference path does), and their gradients come back as column sums of the
first-layer pre-activation gradients (dW_hoisted = colsum (x) input, d input = colsum . W_hoisted).

Numerics follow the reference's fp16 autocast step: fp16 layer inputs / outputs, fp32 accumulation, fp32 activations' derivatives;
table and weight gradients accumulate in fp32 (the reference rounds weight gradients to fp16 and scatters table gradients with fp16
atomics).  Gradients arrive and leave scaled by GradScaler's loss scale like any autocast backward.
"""
import ctypes as C

import numpy as np
import torch
from torch.amp import custom_bwd, custom_fwd

from . import abi
from .frame import GridTable, pack_table

_vp, _u32, _f32, _u64 = C.c_void_p, C.c_uint32, C.c_float, C.c_uint64


class HeadTrainDesc(C.Structure):   # rn_head_train_desc
    _fields_ = [("M", _u32), ("reserved", _u32), ("bound", _f32), ("reserved1", _f32),
                ("xyzs", _vp), ("dirs", _vp), ("grid3d", GridTable), ("grid2d", GridTable),
                ("fwd_blob", _vp), ("bwd_blob", _vp), ("consts", _vp),
                ("sigma", _vp), ("rgb", _vp), ("ambient", _vp), ("sigma_pre", _vp),
                ("acts", _vp), ("dy_dx2", _vp),
                ("d_sigma", _vp), ("d_rgb", _vp), ("d_ambient", _vp),
                ("d_table3", _vp), ("d_table2", _vp), ("d_weights", _vp),
                ("workspace", _vp), ("workspace_bytes", _u64), ("m_valid", _vp)]


abi.register("rn_head_train_acts_bytes", [_u32], _u64)
abi.register("rn_head_train_workspace_bytes", [_u32], _u64)
abi.register("rn_head_train_bwd_blob_bytes", [], _u32)
abi.register("rn_head_train_dw_floats", [], _u32)
abi.register("rn_head_train_forward", [C.POINTER(HeadTrainDesc), _vp])
abi.register("rn_head_train_backward", [C.POINTER(HeadTrainDesc), _vp])
abi.register("rn_pack_grid_table", [_vp, _vp, _vp, _u32, _vp, _vp])
abi.register("rn_pack_head_blobs", [C.POINTER(_vp * 8), _vp, _vp, _vp])
abi.register("rn_head_blob_bytes", [], _u32)

# offsets inside d_weights (csrc/head_train.cuh G_*)
_G = {}
_off = 0
for _name, _shape in (("wa1x", (64, 32)), ("wa2", (64, 64)), ("wa3", (2, 64)), ("ws1", (64, 64)), ("ws2", (64, 64)), ("ws3", (65, 64)),
                      ("wc1", (64, 80)), ("wc2", (3, 64)), ("cs_a1", (64,)), ("cs_s1", (64,)), ("cs_c1", (64,))):
    _G[_name] = (_off, _shape)
    _off += int(np.prod(_shape))
G_FLOATS = _off


def supported(model):
    """the kernels are specialised for the head architecture nerf/network.py builds (see frame.supported)"""
    try:
        return bool(model.encoder.num_levels == 16 and model.encoder.level_dim == 2 and model.encoder.gridtype == 'tiled'
                    and model.encoder.interpolation == 'linear' and not model.encoder.align_corners and model.encoder.input_dim == 3
                    and model.encoder_ambient.num_levels == 16 and model.encoder_ambient.input_dim == 2
                    and model.encoder_ambient.gridtype == 'tiled' and model.encoder_ambient.level_dim == 2
                    and model.hidden_dim == 64 and model.geo_feat_dim == 64 and model.num_layers == 3 and model.num_layers_color == 2
                    and model.hidden_dim_ambient == 64 and model.num_layers_ambient == 3 and model.audio_dim == 64 and model.ambient_dim == 2
                    and model.encoder_dir.degree == 4 and model.individual_dim == 4 and model.exp_eye and model.opt.fp16)
    except AttributeError:
        return False


def _weights(model):
    a, s, c = model.ambient_net.net, model.sigma_net.net, model.color_net.net
    return [a[0].weight, a[1].weight, a[2].weight, s[0].weight, s[1].weight, s[2].weight, c[0].weight, c[1].weight]


class HeadTrainer:
    """buffers of the fused head step, owned by the model (`model._head_trainer`): packed fp16 tables, operand blobs, saved
    activations (928 B/sample, grown on demand), backward workspace, gradient staging"""

    def __init__(self, model):
        L = abi.lib()
        if int(L.rn_head_train_dw_floats()) != G_FLOATS:
            raise RuntimeError("fused_train.py and libradnerf_b200.so disagree about the weight-gradient layout")
        self.dev = model.encoder.embeddings.device
        self.fwd_blob = torch.empty(int(L.rn_head_blob_bytes()), dtype=torch.uint8, device=self.dev)
        self.bwd_blob = torch.empty(int(L.rn_head_train_bwd_blob_bytes()), dtype=torch.uint8, device=self.dev)
        self.tables = {}
        for name, enc in (("t3", model.encoder), ("t2", model.encoder_ambient)):
            packed, first = pack_table(enc)        # allocates the padded layout once (padding rows stay zero)
            self.tables[name] = (packed, first)
        self.consts = torch.zeros(192, device=self.dev)
        self.capacity = 0
        self.d_weights = torch.empty(G_FLOATS, device=self.dev)
        self.packed_for = None

    def refresh(self, model, force=False):
        """re-pack tables and weight blobs when a parameter changed (every optimiser step): three launches"""
        ws = _weights(model)
        tag = tuple((p.data_ptr(), p._version) for p in [model.encoder.embeddings, model.encoder_ambient.embeddings] + ws)
        if tag == self.packed_for and not force:
            return
        L = abi.lib()
        st = abi.cur_stream()
        for name, enc in (("t3", model.encoder), ("t2", model.encoder_ambient)):
            packed, first = self.tables[name]
            abi.check(L.rn_pack_grid_table(abi.ptr(enc.embeddings), abi.ptr(enc.offsets), abi.ptr(first), enc.num_levels, abi.ptr(packed), st))
        arr = (_vp * 8)(*[w.data_ptr() for w in ws])
        abi.check(L.rn_pack_head_blobs(C.byref(arr), abi.ptr(self.fwd_blob), abi.ptr(self.bwd_blob), st))
        self.packed_for = tag

    def ensure(self, M):
        if M <= self.capacity:
            return
        L = abi.lib()
        cap = max(M, int(self.capacity * 1.25))
        cap = (cap + 127) // 128 * 128
        self.acts = torch.empty(int(L.rn_head_train_acts_bytes(cap)), dtype=torch.uint8, device=self.dev)
        self.dy_dx2 = torch.empty(cap, 64, dtype=torch.float16, device=self.dev)
        self.workspace = torch.empty(int(L.rn_head_train_workspace_bytes(cap)), dtype=torch.uint8, device=self.dev)
        self.sigma_pre = torch.empty(cap, device=self.dev)
        self.capacity = cap

    def desc(self, model, xyzs, dirs, M):
        d = HeadTrainDesc()
        d.M, d.bound = M, float(model.bound)
        d.xyzs, d.dirs = xyzs.data_ptr(), (dirs.data_ptr() if dirs is not None else None)
        for field, name, enc in (("grid3d", "t3", model.encoder), ("grid2d", "t2", model.encoder_ambient)):
            packed, first = self.tables[name]
            setattr(d, field, GridTable(packed.data_ptr(), enc.offsets.data_ptr(), float(np.log2(enc.per_level_scale)),
                                        int(enc.base_resolution), first.data_ptr()))
        d.fwd_blob, d.bwd_blob, d.consts = self.fwd_blob.data_ptr(), self.bwd_blob.data_ptr(), self.consts.data_ptr()
        return d


def trainer(model):
    t = getattr(model, "_head_trainer", None)
    if t is None or t.dev != model.encoder.embeddings.device:
        t = model._head_trainer = HeadTrainer(model)
    return t


def _h(t):   # the value an fp16 autocast Linear sees
    return t.detach().half().float()


def _hoisted(model, enc_a, ind_code, eye, out):
    """first-layer products of the per-frame-constant inputs -> out [3*64] (what rn_frame_conditioning computes at inference)"""
    wa1, ws1, wc1 = model.ambient_net.net[0].weight, model.sigma_net.net[0].weight, model.color_net.net[0].weight
    out[0:64] = _h(wa1[:, 32:96]) @ _h(enc_a).reshape(-1)
    out[64:128] = _h(ws1[:, 64]) * _h(eye).reshape(-1)[0]
    out[128:192] = _h(wc1[:, 80:84]) @ _h(ind_code).reshape(-1)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)     # fp32 in, autocast off inside (the kernels do their own fp16)
    def forward(ctx, model, xyzs, dirs, enc_a, ind_code, eye, n_valid, emb3, emb2, wa1, wa2, wa3, ws1, ws2, ws3, wc1, wc2):
        tr = trainer(model)
        xyzs = xyzs.detach().float().contiguous()
        dirs = dirs.detach().float().contiguous()
        M = xyzs.shape[0]
        dev = xyzs.device
        with torch.no_grad():
            tr.refresh(model)
            tr.ensure(M)
            _hoisted(model, enc_a, ind_code, eye, tr.consts)
        # with a device-side sample count the kernels skip the padding rows of the marcher's buffers: their outputs must be defined
        alloc = torch.empty if n_valid is None else torch.zeros
        sigma = alloc(M, device=dev)
        rgb = alloc(M, 3, device=dev)
        ambient = alloc(M, 2, device=dev)
        d = tr.desc(model, xyzs, dirs, M)
        d.sigma, d.rgb, d.ambient, d.sigma_pre = sigma.data_ptr(), rgb.data_ptr(), ambient.data_ptr(), tr.sigma_pre.data_ptr()
        d.acts, d.dy_dx2 = tr.acts.data_ptr(), tr.dy_dx2.data_ptr()
        d.m_valid = None if n_valid is None else n_valid.data_ptr()
        abi.check(abi.lib().rn_head_train_forward(C.byref(d), abi.cur_stream()), "rn_head_train_forward")
        ctx.model, ctx.M, ctx.n_valid = model, M, n_valid
        ctx.save_for_backward(xyzs, dirs, sigma, rgb, ambient, enc_a, ind_code, eye)
        return sigma, rgb, ambient

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, d_sigma, d_rgb, d_ambient):
        model, M = ctx.model, ctx.M
        xyzs, dirs, sigma, rgb, ambient, enc_a, ind_code, eye = ctx.saved_tensors
        tr = trainer(model)
        dev = xyzs.device
        zeros = lambda ref: torch.zeros_like(ref, dtype=torch.float32)     # noqa: E731
        d_sigma = (d_sigma if d_sigma is not None else zeros(sigma)).float().contiguous()
        d_rgb = (d_rgb if d_rgb is not None else zeros(rgb)).float().contiguous()
        d_ambient = (d_ambient if d_ambient is not None else zeros(ambient)).float().contiguous()
        d_t3 = torch.zeros_like(model.encoder.embeddings, dtype=torch.float32)
        d_t2 = torch.zeros_like(model.encoder_ambient.embeddings, dtype=torch.float32)
        d = tr.desc(model, xyzs, dirs, M)
        d.sigma, d.rgb, d.ambient, d.sigma_pre = sigma.data_ptr(), rgb.data_ptr(), ambient.data_ptr(), tr.sigma_pre.data_ptr()
        d.acts, d.dy_dx2 = tr.acts.data_ptr(), tr.dy_dx2.data_ptr()
        d.d_sigma, d.d_rgb, d.d_ambient = d_sigma.data_ptr(), d_rgb.data_ptr(), d_ambient.data_ptr()
        d.d_table3, d.d_table2, d.d_weights = d_t3.data_ptr(), d_t2.data_ptr(), tr.d_weights.data_ptr()
        d.workspace, d.workspace_bytes = tr.workspace.data_ptr(), tr.workspace.numel()
        d.m_valid = None if ctx.n_valid is None else ctx.n_valid.data_ptr()
        abi.check(abi.lib().rn_head_train_backward(C.byref(d), abi.cur_stream()), "rn_head_train_backward")

        def g(name):
            off, shape = _G[name]
            return tr.d_weights[off:off + int(np.prod(shape))].view(shape)
        wa1, ws1, wc1 = model.ambient_net.net[0].weight, model.sigma_net.net[0].weight, model.color_net.net[0].weight
        ea, ey, ic = _h(enc_a).reshape(-1), _h(eye).reshape(-1)[:1], _h(ind_code).reshape(-1)
        cs_a1, cs_s1, cs_c1 = g("cs_a1"), g("cs_s1"), g("cs_c1")
        d_wa1 = torch.cat([g("wa1x"), torch.outer(cs_a1, ea)], 1)
        d_ws1 = torch.cat([g("ws1"), torch.outer(cs_s1, ey)], 1)
        d_wc1 = torch.cat([g("wc1"), torch.outer(cs_c1, ic)], 1)
        d_enc_a = (cs_a1 @ _h(wa1[:, 32:96])).view_as(enc_a).to(enc_a.dtype) if ctx.needs_input_grad[3] else None
        d_ind = (cs_c1 @ _h(wc1[:, 80:84])).view_as(ind_code).to(ind_code.dtype) if ctx.needs_input_grad[4] else None
        d_eye = (cs_s1 @ _h(ws1[:, 64])).view_as(eye).to(eye.dtype) if ctx.needs_input_grad[5] else None
        return (None, None, None, d_enc_a, d_ind, d_eye, None, d_t3, d_t2, d_wa1, g("wa2").clone(), g("wa3").clone(), d_ws1, g("ws2").clone(),
                g("ws3").clone(), d_wc1, g("wc2").clone())


def head_forward(model, xyzs, dirs, enc_a, ind_code, eye, n_valid=None):
    """NeRFNetwork.forward on march_rays_train samples through the fused kernels: -> sigma [M] fp32, color [M,3] fp32 (fp16-rounded
    values, as the autocast path produces), ambient [M,2] fp32.  n_valid: optional int32 DEVICE tensor [1] -- only the first
    min(M, n_valid) rows are samples, the rest is the zero padding of the marcher's buffers (outputs there are zeros, no table
    gradient: the reference evaluates them as points at the origin, all in the same grid cells)."""
    if not supported(model):
        raise NotImplementedError("model configuration outside the fused training kernels' specialisation")
    if enc_a is None or ind_code is None or eye is None:
        raise NotImplementedError("the fused head step needs the audio code, the individual code and the eye value")
    return _HeadFn.apply(model, xyzs, dirs, enc_a, ind_code, eye, n_valid, model.encoder.embeddings, model.encoder_ambient.embeddings, *_weights(model))


@torch.no_grad()
def density(model, xyzs, enc_a, eye):
    """NeRFNetwork.density (nerf/network.py:286-325) for the occupancy update: sigma [M] only, nothing saved, no colour net"""
    tr = trainer(model)
    xyzs = xyzs.float().contiguous()
    M = xyzs.shape[0]
    ind = model.individual_codes[0] if getattr(model, "individual_codes", None) is not None else torch.zeros(4, device=xyzs.device)
    with torch.autocast("cuda", enabled=False):     # the hoisted first-layer terms are fp32 products, as in head_forward
        tr.refresh(model)
        _hoisted(model, enc_a.float(), ind.float(), eye.float(), tr.consts)
    sigma = torch.empty(M, device=xyzs.device)
    d = tr.desc(model, xyzs, None, M)
    d.sigma = sigma.data_ptr()
    abi.check(abi.lib().rn_head_train_forward(C.byref(d), abi.cur_stream()), "rn_head_train_forward (density)")
    return sigma
