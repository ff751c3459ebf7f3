"""Training step of the ray path and its data-parallel gradient exchange ("Training step" row).

`train_step` restates what Trainer.train_step + the optimiser part of Trainer.train_one_epoch do for one batch of rays
(nerf/utils.py:718-808, 1153-1182): render in training mode through the op-by-op path (march_rays_train ->
network -> composite_rays_train, all autograd Functions over the C ABI), per-ray MSE, the alpha-entropy term, the
ambient regulariser outside the face mask, GradScaler backward/step.

`GradSync` is the multi-GPU part: one process per GPU holds a replica and draws its own rays; gradients are summed with
NCCL over NVLink and divided by the world size.  The three hash tables carry ~99% of the gradient bytes (12 MB of
12.1 MB) and their gradients become final at well separated points of the backward pass, so each table is all-reduced
from an autograd hook the moment its gradient is accumulated -- asynchronously, overlapping the remainder of backward --
while every small parameter (MLPs, audio nets, codes) travels in ONE flat bucket once backward is done.  There is no
collective in the forward pass."""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, world=None, table_numel=1 << 18, group=None):
        """params: iterable of parameters to keep in sync; tensors with >= table_numel elements get their own
        asynchronous all-reduce, the rest share a flat bucket."""
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.params = [p for p in params if p.requires_grad]
        self.big = [p for p in self.params if p.numel() >= table_numel]
        self.small = [p for p in self.params if p.numel() < table_numel]
        self.works = []
        self.hooks = []
        self.bytes_last = 0
        if self.world > 1:
            for p in self.big:
                self.hooks.append(p.register_post_accumulate_grad_hook(self._on_table_grad))

    def _on_table_grad(self, p):
        # the gradient of this table is final for this step: start its all-reduce now, backward continues underneath
        self.works.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.bytes_last += p.grad.numel() * p.grad.element_size()

    def finish(self):
        """call after backward(): exchanges the small-parameter bucket, waits for the table reductions, averages"""
        if self.world <= 1:
            return
        grads = [p.grad for p in self.small if p.grad is not None]
        if grads:
            flat = torch.cat([g.reshape(-1).float() for g in grads])
            self.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.bytes_last += flat.numel() * 4
        for w in self.works:
            w.wait()
        self.works = []
        inv = 1.0 / self.world
        if grads:
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g) * inv)
                off += n
        for p in self.big:
            if p.grad is not None:
                p.grad.mul_(inv)

    def begin_step(self):
        self.bytes_last = 0

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def head_loss(out, rgb, face_mask, lambda_amb):
    """nerf/utils.py:749,783-806 for the head phase"""
    loss = ((out["image"] - rgb) ** 2).mean(-1).mean()
    a = out["weights_sum"].clamp(1e-5, 1 - 1e-5)
    loss = loss + 1e-4 * (-a * torch.log2(a) - (1 - a) * torch.log2(1 - a)).mean()
    if lambda_amb:
        loss = loss + lambda_amb * (out["ambient"] * (~face_mask.reshape(-1))).mean()
    return loss


def torso_loss(out, rgb):
    """nerf/utils.py:746-749,787-791 for the torso phase"""
    loss = ((out["torso_color"].view_as(rgb) - rgb) ** 2).mean(-1).mean()
    a = out["torso_alpha"].clamp(1e-5, 1 - 1e-5)
    return loss + 1e-4 * (-a * torch.log2(a) - (1 - a) * torch.log2(1 - a)).mean()


def train_step(model, batch, optimizer, scaler=None, sync=None, lambda_amb=0.1, phase=None):
    """one optimisation step on `batch` = dict(rays_o [1,N,3], rays_d, auds, bg_coords [1,N,2], poses [1,6], eye, index,
    rgb [1,N,3], face_mask [1,N] bool, bg_color [N,3] or scalar).  phase: "head" | "torso" (default: the model's
    --torso flag, as the reference decides).  Returns the detached loss."""
    model.train()
    dev = batch["rays_o"].device
    amp = bool(model.opt.fp16) and dev.type == "cuda"
    if sync is not None:
        sync.begin_step()
    optimizer.zero_grad(set_to_none=False)
    with torch.autocast(dev.type, dtype=torch.float16, enabled=amp):
        out = model.render(batch["rays_o"], batch["rays_d"], batch["auds"], batch["bg_coords"], batch["poses"], eye=batch.get("eye"),
                           index=batch.get("index", 0), bg_color=batch.get("bg_color"), perturb=True, force_all_rays=False,
                           **model.opt.render_kwargs())
        rgb = batch["rgb"]
        if phase is None:
            phase = "torso" if getattr(model, "torso", False) else "head"
        loss = head_loss(out, rgb, batch["face_mask"], lambda_amb) if phase == "head" else torso_loss(out, rgb)
    if scaler is not None and amp:
        scaler.scale(loss).backward()
        if sync is not None:
            sync.finish()
        scaler.step(optimizer)
        scaler.update()
    else:
        loss.backward()
        if sync is not None:
            sync.finish()
        optimizer.step()
    return loss.detach()


def update_extra_state_replicated(model, group=None, share_counters=True):
    """`model.update_extra_state()` on every data-parallel rank such that the replicas stay bit-identical (SURVEY 8(e), rows
    "Occupancy update" and "mean_count").  The update draws a random audio window / pose (Python `random`) and jitters the
    query points inside their cells (`torch.rand_like`, nerf/renderer.py:383-501): rank 0 broadcasts one seed, every rank
    runs the update under it (its own RNG streams -- which feed the per-rank ray sampling -- are restored afterwards), and
    with identical weights, grids and draws all ranks build the same density grid and bitfield without exchanging them.
    share_counters: the (samples, rays) counters of the last <= 16 steps are max-reduced first, so `mean_count`, which sizes
    the marcher's buffers and decides which rays are dropped on overflow, is the same everywhere (and safe for the rank that
    emitted the most samples)."""
    import random
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world <= 1:
        return model.update_extra_state()
    dev = model.step_counter.device
    import os
    seed = torch.tensor([int.from_bytes(os.urandom(4), "little") & 0x7FFFFFFF], dtype=torch.int64, device=dev)   # no RNG stream consumed
    dist.broadcast(seed, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if share_counters:
        dist.all_reduce(model.step_counter, op=dist.ReduceOp.MAX, group=group)
    py_state = random.getstate()
    with torch.random.fork_rng(devices=[dev] if dev.type == "cuda" else []):
        torch.manual_seed(int(seed.item()))
        random.seed(int(seed.item()))
        try:
            return model.update_extra_state()
        finally:
            random.setstate(py_state)


class GraphedTrainStep:
    """`train_step` replayed as ONE CUDA graph per step (single GPU, FusedAdam tail, fp16 autocast).

    Once the marcher sizes its buffers from `mean_count` (after the first occupancy update, raymarching.py:213-229) every
    shape in the step is fixed until the next update, so the whole step -- march, encoders, MLPs, compositing, loss,
    backward, GradScaler inf check, the optimiser sweep and the scaler update -- can be captured once and replayed: one
    launch per step.  What moves between steps stays outside the frozen kernel arguments: the batch is copied into static
    input tensors, the learning rates go through FusedAdam's device-side group table (`publish_groups`), the marcher's
    noise comes from torch's graph-safe generator, the `(samples, rays)` counter row is copied to where the eager step would
    have written it.  A change of `mean_count` (every update_extra_state) or of the batch shape re-captures into the same
    memory pool.  Steps before `mean_count` is known run eagerly, and so does everything if capture fails
    (`self.fallback_reason` says why).

    Status (round 1, one B200, 2^16 rays, ~0.7 M samples): NOT the default.  The capture is correct
    (tests/test_gpu_train.py), but a replay takes 10.3 ms against 8.5 ms for the op-by-op step and each re-capture costs
    25-160 ms (tools/train_bench.py graphed:steady): the step is bound by its ~400 kernels' device work -- a third of it was
    atomic contention in the 2-D grid backward, fixed in csrc/gridencoder_impl.cuh -- not by launch overhead.  It stays as the
    capture-safe scaffolding (static inputs, device-side learning rates, counter bookkeeping) for a fused training step."""

    KEYS = ("rays_o", "rays_d", "auds", "bg_coords", "poses", "eye", "rgb", "face_mask", "bg_color")

    def __init__(self, model, optimizer, scaler, lambda_amb=0.1, phase=None):
        from .optim import FusedAdam
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs radnerf_b200.optim.FusedAdam (its step is capturable: no host reads)")
        self.model, self.opt, self.scaler, self.lambda_amb, self.phase = model, optimizer, scaler, lambda_amb, phase
        self.graph = None
        self.key = None
        self.static = None
        self.loss = None
        self.warm = False
        self.fallback_reason = None
        self.captures = 0
        self.replays = 0
        self.capture_ms = []       # host time of every (re-)capture, synchronisation and instantiation included
        self.pool = None           # every capture allocates from the same private pool: no cudaMalloc after the first one

    def _eager(self, batch):
        self.opt.device_groups = False
        return train_step(self.model, batch, self.opt, self.scaler, None, self.lambda_amb, self.phase)

    def _load(self, batch):
        dev = self.static["rays_o"].device
        for k in self.KEYS:
            if self.static.get(k) is not None:
                self.static[k].copy_(torch.as_tensor(batch[k]), non_blocking=True)
        idx = batch.get("index", 0)
        self.static["index"].copy_(torch.as_tensor(idx, dtype=torch.long).reshape(-1)[:1].to(dev, non_blocking=True))

    def _capture(self, batch):
        import time
        t0 = time.perf_counter()
        m = self.model
        dev = batch["rays_o"].device
        previous = self.graph  # stays alive until the new graph exists, so the shared pool keeps its memory
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        self.static = {k: (torch.as_tensor(batch[k]).to(dev).clone() if batch.get(k) is not None else None) for k in self.KEYS}
        self.static["index"] = torch.zeros(1, dtype=torch.long, device=dev)
        self._load(batch)
        self.opt.publish_groups()
        self.opt.device_groups = True
        local_step = m.local_step
        self.counter_row = local_step % 16
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, pool=self.pool):
                self.loss = train_step(m, self.static, self.opt, self.scaler, None, self.lambda_amb, self.phase)
        finally:
            m.local_step = local_step      # capture ran the host bookkeeping of one step without executing it
        self.graph = g
        del previous
        self.captures += 1
        self.capture_ms.append((time.perf_counter() - t0) * 1e3)

    def __call__(self, batch):
        m = self.model
        amp = bool(m.opt.fp16)
        shapes = tuple((k, tuple(torch.as_tensor(batch[k]).shape)) for k in self.KEYS if batch.get(k) is not None)
        if self.fallback_reason is not None or not amp or m.mean_count <= 0 or not self.warm:
            self.warm = self.warm or m.mean_count > 0      # one eager step in the steady regime: lazy initialisations, grads exist
            return self._eager(batch)
        key = (int(m.mean_count), shapes, self.phase)
        if key != self.key:
            try:
                self._capture(batch)
                self.key = key
            except Exception as e:  # noqa: BLE001
                self.fallback_reason = repr(e)[:300]
                self.graph, self.key = None, None
                torch.cuda.synchronize()
                return self._eager(batch)
        else:
            self._load(batch)
            self.opt.publish_groups()
        self.graph.replay()
        self.replays += 1
        # host bookkeeping of run_cuda's training branch: the counter row of this step, the step index
        row = m.local_step % 16
        if row != self.counter_row:
            m.step_counter[row].copy_(m.step_counter[self.counter_row])
        m.local_step += 1
        return self.loss.detach()
