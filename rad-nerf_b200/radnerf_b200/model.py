"""Host-side mirror of the reference's model/renderer for the hot path (NeRFNetwork + NeRFRenderer).

Why this file exists: the reference's nerf/network.py and nerf/renderer.py run unchanged on top of the drop-in operator
packages in this directory, but they cannot travel to a box without /root/reference, they import seven pip packages that
are absent offline, and they only know the op-by-op execution order.  This module restates the same model:

  * identical parameter / buffer names and shapes, so a reference checkpoint's `model` state-dict loads as is
    (nerf/network.py:91-167, nerf/renderer.py:84-133);
  * `render(rays_o, rays_d, auds, bg_coords, poses, **kw)` with the reference's keyword contract and result keys
    (nerf/renderer.py:158-316, 504-537); `update_extra_state`, `mark_untrained_grid` (renderer.py:318-501);
  * execution: inference = `path="fused"`, one call into the fused sm_100a frame renderer (radnerf_b200.frame), no per-iteration
    host sync; training = the autograd path over the drop-in operators with the head network in the fused training kernels
    (radnerf_b200.fused_train).  The reference's op-by-op inference loop is NOT here: it is a test-side restatement
    (oracle/ops_frame.py, registered through `register_ops_frame`) used by the parity tests and the CPU port; the GPU yardstick is the
    reference's own classes (baseline/stock.py).

`ops` can be swapped for another operator bundle (the CPU oracle, oracle/cpu_backend.py).
"""
import math
import random
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class Options:
    """The subset of the reference's argparse namespace (main.py:12-120, test.py) that reaches the hot path."""
    bound: float = 1.0
    min_near: float = 0.05
    density_thresh: float = 10.0
    density_thresh_torso: float = 0.01
    dt_gamma: float = 1 / 256
    max_steps: int = 16
    exp_eye: bool = True          # -O
    fp16: bool = True             # -O
    torso: bool = False
    smooth_lips: bool = False     # set by --test
    test_train: bool = False
    cuda_ray: bool = True
    train_camera: bool = False
    att: int = 2
    emb: bool = False
    ind_num: int = 10000
    ind_dim: int = 4
    ind_dim_torso: int = 8
    amb_dim: int = 2
    torso_shrink: float = 0.8
    asr_model: str = "cpierse/wav2vec2-large-xlsr-53-esperanto"
    num_rays: int = 4096 * 16
    update_extra_interval: int = 16

    def render_kwargs(self):
        """what `**vars(opt)` contributes to run_cuda's signature (renderer.py:158)"""
        return dict(dt_gamma=self.dt_gamma, max_steps=self.max_steps)


class DefaultOps:
    """operator bundle = this repo's drop-in packages"""

    def __init__(self):
        import raymarching
        from encoding import get_encoder
        from activation import trunc_exp
        self.rm = raymarching
        self.get_encoder = get_encoder
        self.trunc_exp = trunc_exp


# the op-by-op inference frame (a restatement of the reference's host loop) is registered by oracle/ops_frame.py -- tests, smoke() and
# the CPU port only; the product renders frames through radnerf_b200.frame
_OPS_FRAME = None


def register_ops_frame(fn):
    global _OPS_FRAME
    _OPS_FRAME = fn


# ------------------------------------------------------------------------------------------- audio nets / MLP
def _conv_stack(chans, stride):
    layers = []
    for cin, cout in zip(chans[:-1], chans[1:]):
        layers += [nn.Conv1d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
    return nn.Sequential(*layers)


class AudioAttNet(nn.Module):
    """attention over the 8-frame window (network.py:10-37): 5 x Conv1d(k3) 64->16->8->4->2->1, Linear(8,8)+softmax."""

    def __init__(self, dim_aud=64, seq_len=8):
        super().__init__()
        self.seq_len, self.dim_aud = seq_len, dim_aud
        self.attentionConvNet = _conv_stack([dim_aud, 16, 8, 4, 2, 1], stride=1)
        self.attentionNet = nn.Sequential(nn.Linear(seq_len, seq_len, bias=True), nn.Softmax(dim=1))

    def forward(self, x):  # x [1, seq_len, dim_aud]
        y = self.attentionConvNet(x.permute(0, 2, 1))
        y = self.attentionNet(y.view(1, self.seq_len)).view(1, self.seq_len, 1)
        return torch.sum(y * x, dim=1)


class AudioNet(nn.Module):
    """per-frame audio feature CNN (network.py:41-67): 4 x Conv1d(k3,s2) dim_in->32->32->64->64 over a 16-wide window,
    then Linear 64->64->dim_aud with LeakyReLU(0.02)."""

    def __init__(self, dim_in=29, dim_aud=64, win_size=16):
        super().__init__()
        self.win_size, self.dim_aud = win_size, dim_aud
        self.encoder_conv = _conv_stack([dim_in, 32, 32, 64, 64], stride=2)
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, dim_aud))

    def forward(self, x):
        half_w = int(self.win_size / 2)
        x = x[:, :, 8 - half_w:8 + half_w]
        return self.encoder_fc1(self.encoder_conv(x).squeeze(-1))


class MLP(nn.Module):
    """bias-free Linear stack with ReLU between layers (network.py:69-88)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.net = nn.ModuleList([
            nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=False)
            for l in range(num_layers)])

    def forward(self, x):
        for l, layer in enumerate(self.net):
            x = layer(x)
            if l != self.num_layers - 1:
                x = F.relu(x, inplace=True)
        return x


# ------------------------------------------------------------------------------------------- the model
class NeRFNetwork(nn.Module):
    def __init__(self, opt: Options = None, ops=None, num_layers=3, hidden_dim=64, geo_feat_dim=64, num_layers_color=2,
                 hidden_dim_color=64, audio_dim=64, num_layers_ambient=3, hidden_dim_ambient=64, ambient_dim=2):
        super().__init__()
        self.opt = opt = opt or Options()
        self.ops = ops or DefaultOps()
        get_encoder = self.ops.get_encoder

        # ---- renderer state (renderer.py:63-133)
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.density_scale = 1
        self.min_near, self.density_thresh, self.density_thresh_torso = opt.min_near, opt.density_thresh, opt.density_thresh_torso
        self.exp_eye, self.test_train, self.smooth_lips = opt.exp_eye, opt.test_train, opt.smooth_lips
        self.torso, self.cuda_ray, self.train_camera = opt.torso, opt.cuda_ray, opt.train_camera
        b = opt.bound
        aabb = torch.tensor([-b, -b / 2, -b, b, b / 2, b], dtype=torch.float32)
        self.register_buffer('aabb_train', aabb)
        self.register_buffer('aabb_infer', aabb.clone())
        self.individual_num, self.individual_dim = opt.ind_num, opt.ind_dim
        if self.individual_dim > 0:
            self.individual_codes = nn.Parameter(torch.randn(self.individual_num, self.individual_dim) * 0.1)
        if self.torso:
            self.individual_dim_torso = opt.ind_dim_torso
            if self.individual_dim_torso > 0:
                self.individual_codes_torso = nn.Parameter(torch.randn(self.individual_num, self.individual_dim_torso) * 0.1)
        if self.train_camera:
            self.camera_dR = nn.Parameter(torch.zeros(self.individual_num, 3))
            self.camera_dT = nn.Parameter(torch.zeros(self.individual_num, 3))
        self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
        self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
        self.mean_density = 0
        self.iter_density = 0
        if self.torso:
            self.register_buffer('density_grid_torso', torch.zeros([self.grid_size ** 2]))
        self.mean_density_torso = 0
        self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
        self.mean_count = 0
        self.local_step = 0
        if self.smooth_lips:
            self.enc_a = None

        # ---- networks (network.py:111-167)
        self.emb = opt.emb
        if 'esperanto' in opt.asr_model:
            self.audio_in_dim = 44
        elif 'deepspeech' in opt.asr_model:
            self.audio_in_dim = 29
        else:
            self.audio_in_dim = 32
        if self.emb:
            self.embedding = nn.Embedding(self.audio_in_dim, self.audio_in_dim)
        self.audio_dim = audio_dim
        self.audio_net = AudioNet(self.audio_in_dim, self.audio_dim)
        self.att = opt.att
        if self.att > 0:
            self.audio_att_net = AudioAttNet(self.audio_dim)

        grid = dict(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, interpolation='linear')
        self.encoder, self.in_dim = get_encoder('tiledgrid', input_dim=3, desired_resolution=2048 * self.bound, **grid)
        self.encoder_ambient, self.in_dim_ambient = get_encoder('tiledgrid', input_dim=ambient_dim, desired_resolution=2048, **grid)
        self.num_layers_ambient, self.hidden_dim_ambient, self.ambient_dim = num_layers_ambient, hidden_dim_ambient, ambient_dim
        self.ambient_net = MLP(self.in_dim + self.audio_dim, self.ambient_dim, self.hidden_dim_ambient, self.num_layers_ambient)

        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.eye_dim = 1 if self.exp_eye else 0
        self.sigma_net = MLP(self.in_dim + self.in_dim_ambient + self.eye_dim, 1 + self.geo_feat_dim, self.hidden_dim, self.num_layers)

        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.encoder_dir, self.in_dim_dir = get_encoder('spherical_harmonics')
        self.color_net = MLP(self.in_dim_dir + self.geo_feat_dim + self.individual_dim, 3, self.hidden_dim_color, self.num_layers_color)

        if self.torso:
            self.torso_deform_encoder, self.torso_deform_in_dim = get_encoder('frequency', input_dim=2, multires=10)
            self.pose_encoder, self.pose_in_dim = get_encoder('frequency', input_dim=6, multires=4)
            self.torso_deform_net = MLP(self.torso_deform_in_dim + self.pose_in_dim + self.individual_dim_torso, 2, 64, 3)
            self.torso_encoder, self.torso_in_dim = get_encoder('tiledgrid', input_dim=2, desired_resolution=2048, **grid)
            self.torso_net = MLP(self.torso_in_dim + self.torso_deform_in_dim + self.pose_in_dim + self.individual_dim_torso, 4, 32, 3)

        # data-dependent statistics the occupancy update samples from (set by the data provider in the reference)
        self.aud_features = None  # [n, dim, 16]
        self.eye_area = None      # [n, 1]
        self.poses = None         # [n, 4, 4]

    # ------------------------------------------------------------------------------------- network pieces
    # (the op-by-op network: autograd fallback of the training step, the torso branch, the parity yardstick of the fused kernels.
    #  Every per-frame vector -- audio code, eye value, individual codes, encoded pose -- is broadcast over the batch with _rows.)
    @staticmethod
    def _rows(v, n):
        """a per-frame vector [1, k] (or [k]) as n identical rows"""
        return v.reshape(1, -1).expand(n, -1)

    def encode_audio(self, a):
        """[8, dim, 16] window -> [1, 64] (network.py:170-185); None passes through (no audio conditioning)"""
        if a is None:
            return None
        feats = self.embedding(a).transpose(-1, -2).contiguous() if self.emb else a
        code = self.audio_net(feats)
        return self.audio_att_net(code.unsqueeze(0)) if self.att > 0 else code

    def _ambient_and_features(self, x, enc_a):
        """(ambient coordinate [N,2] in [-1,1], spatial features, ambient-grid features) -- network.py:233-247"""
        n = x.shape[0]
        spatial = self.encoder(x, bound=self.bound)
        if enc_a is None:
            ambient = x.new_zeros(n, self.ambient_dim)
        else:
            ambient = torch.tanh(self.ambient_net(torch.cat([spatial, self._rows(enc_a, n)], 1)).float())
        return ambient, spatial, self.encoder_ambient(ambient, bound=1)

    def _sigma_head(self, spatial, ambient_feat, e, n):
        """density MLP on [spatial | ambient | eye] -> (sigma [N], geo_feat [N,64]) -- network.py:249-262"""
        cols = [spatial, ambient_feat] + ([self._rows(e, n)] if e is not None else [])
        out = self.sigma_net(torch.cat(cols, -1))
        return self.ops.trunc_exp(out[..., 0]), out[..., 1:]

    def forward(self, x, d, enc_a, c, e=None):
        """x [N,3] in [-bound,bound]; d [N,3]; enc_a [1,64]; c [ind_dim]; e [1,1] -> sigma [N], color [N,3], ambient [N,2]
        (network.py:222-283)"""
        n = x.shape[0]
        ambient, spatial, ambient_feat = self._ambient_and_features(x, enc_a)
        sigma, geo_feat = self._sigma_head(spatial, ambient_feat, e, n)
        cols = [self.encoder_dir(d), geo_feat] + ([self._rows(c, n)] if c is not None else [])
        return sigma, torch.sigmoid(self.color_net(torch.cat(cols, -1))), ambient

    def density(self, x, enc_a, e=None):
        """(network.py:286-325)"""
        _, spatial, ambient_feat = self._ambient_and_features(x, enc_a)
        sigma, geo_feat = self._sigma_head(spatial, ambient_feat, e, x.shape[0])
        return {'sigma': sigma, 'geo_feat': geo_feat}

    def forward_torso(self, x, poses, enc_a, c=None):
        """x [N,2] in [-1,1]; poses [1,6] -> alpha [N,1], color [N,3], dx [N,2] (network.py:188-219): a pose-conditioned deformation of
        the pixel coordinate, then the 2-D torso grid + MLP at the deformed position"""
        n = x.shape[0]
        pix = x * self.opt.torso_shrink
        cond = [self.torso_deform_encoder(pix), self._rows(self.pose_encoder(poses), n)] + ([self._rows(c, n)] if c is not None else [])
        cond = torch.cat(cond, -1)
        offset = self.torso_deform_net(cond)
        feat = self.torso_encoder((pix + offset).clamp(-1, 1), bound=1)
        out = torch.sigmoid(self.torso_net(torch.cat([feat, cond], -1)))
        return out[..., :1], out[..., 1:], offset

    def get_params(self, lr, lr_net, wd=0):
        """optimizer groups (network.py:329-361)"""
        if self.torso:
            params = [{'params': self.torso_encoder.parameters(), 'lr': lr},
                      {'params': self.torso_net.parameters(), 'lr': lr_net, 'weight_decay': wd},
                      {'params': self.torso_deform_net.parameters(), 'lr': lr_net, 'weight_decay': wd}]
            if self.individual_dim_torso > 0:
                params.append({'params': self.individual_codes_torso, 'lr': lr_net, 'weight_decay': wd})
            return params
        params = [{'params': self.audio_net.parameters(), 'lr': lr_net, 'weight_decay': wd},
                  {'params': self.encoder.parameters(), 'lr': lr},
                  {'params': self.encoder_ambient.parameters(), 'lr': lr},
                  {'params': self.ambient_net.parameters(), 'lr': lr_net, 'weight_decay': wd},
                  {'params': self.sigma_net.parameters(), 'lr': lr_net, 'weight_decay': wd},
                  {'params': self.color_net.parameters(), 'lr': lr_net, 'weight_decay': wd}]
        if self.att > 0:
            params.append({'params': self.audio_att_net.parameters(), 'lr': lr_net * 5, 'weight_decay': wd})
        if self.emb:
            params.append({'params': self.embedding.parameters(), 'lr': lr})
        if self.individual_dim > 0:
            params.append({'params': self.individual_codes, 'lr': lr_net, 'weight_decay': wd})
        if self.train_camera:
            params.append({'params': self.camera_dT, 'lr': 1e-5, 'weight_decay': 0})
            params.append({'params': self.camera_dR, 'lr': 1e-5, 'weight_decay': 0})
        return params

    # ------------------------------------------------------------------------------------- renderer
    def reset_extra_state(self):
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def _frame_conditioning(self, auds, index):
        """per-frame constants: the audio code -- exponentially smoothed over frames when lip smoothing is on (0.35 of the previous
        frame's code) -- and the individual code of the frame (renderer.py:187-204)"""
        code = self.encode_audio(auds)
        if code is not None and self.smooth_lips:
            previous = self.enc_a
            if previous is not None:
                code = 0.35 * previous + (1 - 0.35) * code
            self.enc_a = code
        individual = None
        if self.individual_dim > 0:
            individual = self.individual_codes[index if self.training else 0]
        return code, individual

    def run_cuda(self, rays_o, rays_d, auds, bg_coords, poses, eye=None, index=0, dt_gamma=0, bg_color=None, perturb=False,
                 force_all_rays=False, max_steps=1024, T_thresh=1e-4, **kwargs):
        """one batch of rays through the op-by-op operators with the reference's result keys (renderer.py:158-316).

        Training mode is the product's autograd path: marcher -> head network (fused kernels when the step is the one they
        implement) -> compositor.  The op-by-op INFERENCE loop of the reference (renderer.py:229-262: march / network / composite per
        iteration with a host synchronisation each) is not part of the product -- inference is `path="fused"` -- and lives with the rest
        of the reference restatements in oracle/ops_frame.py, which registers itself here for the parity tests and the CPU port."""
        dev = rays_o.device
        lead = rays_o.shape[:-1]
        origins, directions = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        pixels = bg_coords.contiguous().view(-1, 2)
        n_rays = origins.shape[0]
        box = self.aabb_train if self.training else self.aabb_infer
        nears, fars = (t.detach() for t in self.ops.rm.near_far_from_aabb(origins, directions, box, self.min_near))
        enc_a, ind_code = self._frame_conditioning(auds, index)
        out = {}
        if self.training:
            weights_sum, depth, image = self._training_batch(out, origins, directions, nears, fars, enc_a, ind_code, eye, dt_gamma, perturb,
                                                             force_all_rays, max_steps)
        else:
            if _OPS_FRAME is None:
                raise RuntimeError("the op-by-op inference frame is test infrastructure (oracle/ops_frame.py registers it); "
                                   "render(..., path='fused') is the inference path of this library")
            weights_sum, depth, image = _OPS_FRAME(self, origins, directions, nears, fars, enc_a, ind_code, eye, dt_gamma, perturb, max_steps, T_thresh)
        background = 1 if bg_color is None else bg_color
        if self.torso:
            background = self._torso_over(out, pixels, poses, enc_a, index, background, n_rays, dev)
        out['depth'] = (torch.clamp(depth - nears, min=0) / (fars - nears)).view(*lead)
        out['image'] = (image + (1 - weights_sum).unsqueeze(-1) * background).view(*lead, 3).clamp(0, 1)
        return out

    def _training_batch(self, out, origins, directions, nears, fars, enc_a, ind_code, eye, dt_gamma, perturb, force_all_rays, max_steps):
        """renderer.py:207-236: all samples of all rays at once, sized by the running estimate `mean_count`"""
        rm = self.ops.rm
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        xyzs, dirs, deltas, rays = rm.march_rays_train(origins, directions, self.bound, self.density_bitfield, self.cascade, self.grid_size,
                                                       nears, fars, counter, self.mean_count, perturb, 128, force_all_rays, dt_gamma, max_steps)
        if self._use_fused_train(enc_a, ind_code, eye):
            from . import fused_train    # ONE forward kernel (+ 2 backward kernels) instead of 8 GEMMs + ~150 elementwise launches
            # samples occupy the first min(counter, budget) rows of the buffers; the fused kernels skip the zero padding behind them
            n_valid = counter[:1]
            budget = getattr(getattr(rm, "raymarching", None), "sample_budget", None)
            if budget is not None and budget.active is not None:
                n_valid = torch.minimum(n_valid, budget.active.budget)
            sigmas, rgbs, ambient = fused_train.head_forward(self, xyzs, dirs, enc_a, ind_code, eye, n_valid=n_valid)
        else:
            sigmas, rgbs, ambient = self(xyzs, dirs, enc_a, ind_code, eye)
        weights_sum, ambient_sum, depth, image = rm.composite_rays_train(self.density_scale * sigmas, rgbs, ambient.abs().sum(-1), deltas, rays)
        out['weights_sum'], out['ambient'] = weights_sum, ambient_sum
        return weights_sum, depth, image

    def _torso_over(self, out, pixels, poses, enc_a, index, background, n_rays, dev):
        """the 2-D torso layer composited over the background colour (renderer.py:267-297): evaluated only where the 2-D occupancy
        grid, sampled bilinearly at the pixel's background coordinate, exceeds its threshold"""
        code = None
        if self.individual_dim_torso > 0:
            code = self.individual_codes_torso[index] if self.training else self.individual_codes_torso[0]
        occupancy = F.grid_sample(self.density_grid_torso.view(1, 1, self.grid_size, self.grid_size), pixels.view(1, -1, 1, 2),
                                  align_corners=True).view(-1)
        hit = occupancy > min(self.density_thresh_torso, self.mean_density_torso)
        alpha, color = torch.zeros([n_rays, 1], device=dev), torch.zeros([n_rays, 3], device=dev)
        if hit.any():
            a, c, deform = self.forward_torso(pixels[hit], poses, enc_a, code)
            alpha[hit], color[hit] = a.float(), c.float()
            out['deform'] = deform
        over = color * alpha + background * (1 - alpha)
        out['torso_alpha'], out['torso_color'] = alpha, over
        return over

    # ---- mean_density / mean_count: plain Python numbers for every reader (the reference keeps them as such, renderer.py:470, 493), but
    #      update_extra_state leaves them on the DEVICE and the host copy is only made when somebody asks: no host synchronisation
    #      inside the training loop (GraphedTrainStep consumes the device values directly)
    @property
    def mean_density(self):
        if self._mean_density_host is None:
            self._mean_density_host = float(self._mean_density_dev.item())
        return self._mean_density_host

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density_host, self._mean_density_dev = float(v), None

    @property
    def mean_count(self):
        if self._mean_count_host is None:
            self._mean_count_host = int(self._mean_count_dev.item())
        return self._mean_count_host

    @mean_count.setter
    def mean_count(self, v):
        self._mean_count_host, self._mean_count_dev = int(v), None
        self.mean_count_version = getattr(self, "mean_count_version", 0) + 1

    # fused_train: "auto" (default) = use the fused training kernels of radnerf_b200.fused_train whenever the step is the one they
    # implement (CUDA, fp16 autocast, the stock head architecture, this repo's operators); False = always op by op (the parity yardstick)
    fused_train = "auto"

    def _use_fused_train(self, enc_a, ind_code, eye):
        if not self.fused_train or not isinstance(self.ops, DefaultOps) or self.density_bitfield.device.type != "cuda":
            return False
        if enc_a is None or ind_code is None or eye is None or not torch.is_autocast_enabled():
            return False
        from . import fused_train
        return fused_train.supported(self)

    def render(self, rays_o, rays_d, auds, bg_coords, poses, staged=False, max_ray_batch=4096, path=None, **kwargs):
        """entry point with the reference's contract (renderer.py:504-537); cuda_ray never stages.
        path: "fused" = the whole frame inside the fused sm_100a renderer (inference); "ops" = operator by operator (training always;
        inference only with the test-side loop of oracle/ops_frame.py registered); None = fused for inference whenever the fused
        kernels implement this model on this device, "ops" otherwise."""
        if not self.training and path != "ops":
            from . import frame
            if path == "fused" or (rays_o.is_cuda and isinstance(self.ops, DefaultOps) and frame.supported(self) and _OPS_FRAME is None):
                return frame.render_frame(self, rays_o, rays_d, auds, bg_coords, poses, **kwargs)
        return self.run_cuda(rays_o, rays_d, auds, bg_coords, poses, **kwargs)

    # ------------------------------------------------------------------------------------- occupancy maintenance
    def _cell_centres(self):
        """[H^3, 3] cell coordinates mapped to [-1, 1], listed in MORTON order: row n belongs to row n of `density_grid`.
        Built once with morton3D_invert, so the maintenance passes below write whole grid rows instead of scattering."""
        dev = self.density_bitfield.device
        cache = getattr(self, "_cells", None)
        if cache is None or cache.device != dev:
            n = self.grid_size ** 3
            if hasattr(self.ops.rm, "morton3D_invert"):
                ijk = self.ops.rm.morton3D_invert(torch.arange(n, dtype=torch.int32, device=dev))
            else:   # operator bundles without the inverse (CPU port): invert the forward map once
                g = torch.arange(self.grid_size, dtype=torch.int32, device=dev)
                ijk_lin = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
                ijk = torch.empty_like(ijk_lin)
                ijk[self.ops.rm.morton3D(ijk_lin).long()] = ijk_lin
            cache = self._cells = 2 * ijk.float() / (self.grid_size - 1) - 1
        return cache

    def _cascade_geometry(self, cas):
        bound = min(2 ** cas, self.bound)
        half_cell = bound / self.grid_size
        return bound - half_cell, half_cell   # scale applied to the [-1,1] centres, half a cell in world units

    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """cells that no training camera sees get density -1 (renderer.py:318-381: in front of the camera and inside the
        frustum widened by one cell).  All cells of a cascade are tested at once, cameras in chunks of S."""
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        fx, fy, cx, cy = [float(v) for v in intrinsic]
        dev = self.density_bitfield.device
        poses = poses.to(dev).float()
        rot, origin = poses[:, :3, :3], poses[:, :3, 3]
        centres = self._cell_centres()
        block = 1 << 18   # cells per pass: bounds the [S, block, 3] intermediate
        seen = torch.zeros_like(self.density_grid, dtype=torch.bool)
        for cas in range(self.cascade):
            scale, half_cell = self._cascade_geometry(cas)
            for c0 in range(0, centres.shape[0], block):
                world = centres[c0:c0 + block] * scale
                for p0 in range(0, poses.shape[0], S):
                    cam = (world.unsqueeze(0) - origin[p0:p0 + S].unsqueeze(1)) @ rot[p0:p0 + S]     # world -> camera, [S, n, 3]
                    depth = cam[..., 2]
                    inside = (depth > 0) & (cam[..., 0].abs() < cx / fx * depth + half_cell * 2) & \
                             (cam[..., 1].abs() < cy / fy * depth + half_cell * 2)
                    seen[cas, c0:c0 + block] |= inside.any(0)
        self.density_grid[~seen] = -1

    def _query_density_grid(self, enc_a, eye, chunk=1 << 19):
        """density of every cell of every cascade at a jittered position inside the cell -> [cascade, H^3] (Morton order)"""
        centres = self._cell_centres()
        fresh = torch.empty_like(self.density_grid)
        for cas in range(self.cascade):
            scale, half_cell = self._cascade_geometry(cas)
            for c0 in range(0, centres.shape[0], chunk):
                pts = centres[c0:c0 + chunk] * scale
                pts += (torch.rand_like(pts) * 2 - 1) * half_cell
                if self._use_fused_train(enc_a, 0, eye):
                    from . import fused_train
                    sigma = fused_train.density(self, pts, enc_a, eye)     # sigma-only pass of the fused forward kernel
                else:
                    sigma = self.density(pts, enc_a, eye)['sigma'].reshape(-1).detach()
                fresh[cas, c0:c0 + chunk] = sigma.to(fresh.dtype) * self.density_scale
        return fresh

    def _query_torso_grid(self, pose6, enc_a, ind_code):
        """torso alpha of every cell of the 2-D grid at a jittered position; stored x/y-transposed as the reference does"""
        H = self.grid_size
        g = torch.arange(H, dtype=torch.float32, device=self.density_bitfield.device)
        gx, gy = torch.meshgrid(g, g, indexing='ij')
        xy = torch.stack([gx.reshape(-1), gy.reshape(-1)], -1)             # cell (x, y), x-major
        half_cell = 1 / H
        pts = (2 * xy / (H - 1) - 1) * (1 - half_cell)
        pts += (torch.rand_like(pts) * 2 - 1) * half_cell
        alphas, _, _ = self.forward_torso(pts, pose6, enc_a, ind_code)
        out = torch.zeros_like(self.density_grid_torso)
        out[(xy[:, 1] * H + xy[:, 0]).long()] = alphas.squeeze(1).float()     # row = y * H + x
        return out

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """occupancy maintenance between epochs (renderer.py:383-501): re-query the density at every cell for a random audio
        window, dilate, merge into the running grid by max(decayed old, fresh) where both are valid, rebuild the bitfield
        with threshold min(mean density, density_thresh); in the torso phase the same for the 2-D alpha grid (5x5 max-pool
        as dilation); finally the average sample count of the last <= 16 training steps."""
        rm = self.ops.rm
        dev = self.density_bitfield.device
        rand_idx = random.randint(0, self.aud_features.shape[0] - 1)
        from .synthetic import audio_window
        feats = np.asarray(self.aud_features.cpu()) if torch.is_tensor(self.aud_features) else self.aud_features
        enc_a = self.encode_audio(torch.as_tensor(audio_window(feats, rand_idx, self.att)).to(dev))

        if not self.torso:   # the head grid is frozen while the torso trains
            eye = self.eye_area[[rand_idx]].to(dev) if self.exp_eye else None
            fresh = rm.morton3D_dilation(self._query_density_grid(enc_a, eye))
            if isinstance(self.ops, DefaultOps) and dev.type == "cuda":
                # merge, mean and re-pack on the device: two launches, no boolean-mask indexing (a hidden sync each), no `.item()`
                from . import abi
                if getattr(self, "_occ_ws", None) is None or self._occ_ws.device != dev:
                    self._occ_ws = torch.zeros(int(abi.lib().rn_occupancy_merge_workspace_bytes()) // 8 + 1, dtype=torch.float64, device=dev)
                mean = torch.empty(1, dtype=torch.float32, device=dev)
                abi.call("rn_occupancy_merge", self.density_grid, fresh.contiguous(), self.density_grid.numel(), float(decay), self._occ_ws, mean)
                abi.call("rn_packbits_min", self.density_grid, self.density_bitfield.numel(), float(self.density_thresh), mean, self.density_bitfield)
                torch.autograd.graph.increment_version(self.density_bitfield)     # written through the raw pointer (see PackbitsFn)
                self._mean_density_dev, self._mean_density_host = mean, None     # untrained (-1) cells count as empty
                self.iter_density += 1
            else:
                both_valid = (self.density_grid >= 0) & (fresh >= 0)
                self.density_grid[both_valid] = torch.maximum(self.density_grid[both_valid] * decay, fresh[both_valid])
                self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()   # untrained (-1) cells count as empty
                self.iter_density += 1
                self.density_bitfield = rm.packbits(self.density_grid, min(self.mean_density, self.density_thresh), self.density_bitfield)
        else:
            from .posemath import convert_poses
            k = random.randint(0, self.poses.shape[0] - 1)
            ind_code = self.individual_codes_torso[[k]] if self.opt.ind_dim_torso > 0 else None
            fresh = self._query_torso_grid(convert_poses(self.poses[[k]]).to(dev), enc_a, ind_code)
            fresh = F.max_pool2d(fresh.view(1, 1, self.grid_size, self.grid_size), kernel_size=5, stride=1, padding=2).view(-1)
            self.density_grid_torso = torch.maximum(self.density_grid_torso * decay, fresh)
            self.mean_density_torso = torch.mean(self.density_grid_torso).item()

        steps = min(16, self.local_step)
        if steps > 0:
            if dev.type == "cuda":      # int(sum / steps) of the reference, left on the device; read back lazily (`mean_count` property)
                self._mean_count_dev = torch.div(self.step_counter[:steps, 0].sum(), steps, rounding_mode="floor").to(torch.int32).reshape(1)
                self._mean_count_host = None
                self.mean_count_version = getattr(self, "mean_count_version", 0) + 1
            else:
                self.mean_count = int(self.step_counter[:steps, 0].sum().item() / steps)
        self.local_step = 0
