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
import os

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, world=None, table_numel=1 << 18, group=None, overlap=True):
        """params: iterable of parameters to keep in sync; tensors with >= table_numel elements get their own all-reduce
        (asynchronous, from an autograd hook, when `overlap`), the rest share a flat bucket."""
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.params = [p for p in params if p.requires_grad]
        self.big = [p for p in self.params if p.numel() >= table_numel]
        self.small = [p for p in self.params if p.numel() < table_numel]
        self.works = []
        self.hooks = []
        self.bytes_last = 0
        self.overlap = bool(overlap)
        self.flat = None            # flatten(): the small parameters' gradients as views of ONE buffer
        self.flat_params = []
        if self.world > 1:
            for p in self.big:
                self.hooks.append(p.register_post_accumulate_grad_hook(self._on_table_grad))

    def _avg_native(self):
        # NCCL averages inside the collective; gloo (the CPU tests) only sums
        return dist.get_backend(self.group) == "nccl"

    def _on_table_grad(self, p):
        # the gradient of this table is final for this step: start its all-reduce now, backward continues underneath
        if not self.overlap or (p.is_cuda and torch.cuda.is_current_stream_capturing()):
            return
        op = dist.ReduceOp.AVG if self._avg_native() else dist.ReduceOp.SUM
        self.works.append((dist.all_reduce(p.grad, op=op, group=self.group, async_op=True), p))
        self.bytes_last += p.grad.numel() * p.grad.element_size()

    @torch.no_grad()
    def flatten(self):
        """re-point the gradients of the small parameters THAT HAVE ONE (call after a first backward) at slices of one flat fp32
        buffer, so that the bucket is exchanged in place: no gather before and no scatter after the all-reduce (about 90 small
        launches per step otherwise).  Parameters without a gradient keep `grad is None` (an optimiser skips them, as torch's does)."""
        ps = [p for p in self.small if p.grad is not None and p.grad.dtype == torch.float32]
        if not ps:
            return
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            n = p.numel()
            view = flat[off:off + n].view_as(p)
            view.copy_(p.grad)
            p.grad = view
            off += n
        self.flat, self.flat_params = flat, ps

    def finish(self):
        """call after backward(): exchanges the small-parameter bucket, waits for the table reductions, averages"""
        if self.world <= 1:
            return
        avg = self._avg_native()
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        inv = 1.0 / self.world
        started = {id(p) for _, p in self.works}
        late = [p for p in self.big if p.grad is not None and id(p) not in started]     # no hook fired (overlap off / captured backward)
        loose = [p for p in self.small if p.grad is not None and not any(p is q for q in self.flat_params)]
        works = []
        for p in late:
            works.append((dist.all_reduce(p.grad, op=op, group=self.group, async_op=True), p if not avg else None))
            self.bytes_last += p.grad.numel() * p.grad.element_size()
        if self.flat is not None:
            works.append((dist.all_reduce(self.flat, op=op, group=self.group, async_op=True), self.flat if not avg else None))
            self.bytes_last += self.flat.numel() * 4
        bucket = None
        if loose:
            bucket = torch.cat([p.grad.reshape(-1).float() for p in loose])
            works.append((dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True), None))
            self.bytes_last += bucket.numel() * 4
        for w, p in self.works:       # hook-driven table reductions
            w.wait()
            if not avg:
                p.grad.mul_(inv)
        for w, t in works:
            w.wait()
            if t is not None:
                (t.grad if isinstance(t, torch.nn.Parameter) else t).mul_(inv)
        self.works = []
        if bucket is not None:
            off = 0
            for p in loose:
                n = p.numel()
                p.grad.copy_(bucket[off:off + n].view_as(p.grad) * inv)
                off += n

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


def forward_backward(model, batch, optimizer, scaler=None, lambda_amb=0.1, phase=None):
    """first half of a step: zero_grad, render in training mode, loss, (scaled) backward.  Returns the loss tensor."""
    model.train()
    dev = batch["rays_o"].device
    amp = bool(model.opt.fp16) and dev.type == "cuda"
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
    else:
        loss.backward()
    return loss


def optimizer_tail(model, batch, optimizer, scaler=None):
    """second half: GradScaler's inf check + the optimiser step (one sweep with FusedAdam) + the scale update"""
    amp = bool(model.opt.fp16) and batch["rays_o"].device.type == "cuda"
    if scaler is not None and amp:
        scaler.step(optimizer)
        scaler.update()
    else:
        optimizer.step()


def train_step(model, batch, optimizer, scaler=None, sync=None, lambda_amb=0.1, phase=None):
    """one optimisation step on `batch` = dict(rays_o [1,N,3], rays_d, auds, bg_coords [1,N,2], poses [1,6], eye, index,
    rgb [1,N,3], face_mask [1,N] bool, bg_color [N,3] or scalar).  phase: "head" | "torso" (default: the model's
    --torso flag, as the reference decides).  Returns the detached loss."""
    if sync is not None:
        sync.begin_step()
    loss = forward_backward(model, batch, optimizer, scaler, lambda_amb, phase)
    if sync is not None:
        sync.finish()
    optimizer_tail(model, batch, optimizer, scaler)
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
    """`train_step` replayed from CUDA graphs (FusedAdam tail, fp16 autocast): one launch per step on one GPU, two around the
    gradient exchange when data-parallel.

    A steady-state step of the fused head path is ~1.5 ms of device work issued as ~350 launches that cost the host ~5 ms
    (tools/train_timeline.py): it is launch-bound, so it is captured once and replayed.  What moves between steps stays outside
    the frozen kernel arguments: the batch is copied into static input tensors, the learning rates go through FusedAdam's
    device-side group table (`publish_groups`), the marcher's noise comes from torch's graph-safe generator, the `(samples, rays)`
    counter row is copied to where the eager step would have written it, and the marcher's sample budget -- the reference's running
    estimate `mean_count`, which `update_extra_state` changes every 16 steps (renderer.py:489-493) -- is a DEVICE scalar
    (raymarching.sample_budget / rn_march_rays_train_budget): the sample buffers keep a fixed capacity (mean_count rounded up to
    `bucket` slots) and the same rays are kept / dropped as with buffers of exactly mean_count slots.  Only a change of the capacity
    bucket or of the batch shape re-captures (into the same memory pool).

    Data-parallel (`sync` = a GradSync): graph A = zero_grad .. backward, then the NCCL all-reduces are issued eagerly (three
    collectives: two tables + the flat bucket of `sync.flatten()`, exchanged in place), then graph B = inf check + optimiser sweep +
    scale update.  Steps before `mean_count` is known run eagerly, and so does everything if capture fails
    (`self.fallback_reason` says why)."""

    KEYS = ("rays_o", "rays_d", "auds", "bg_coords", "poses", "eye", "rgb", "face_mask", "bg_color")

    def __init__(self, model, optimizer, scaler, lambda_amb=0.1, phase=None, sync=None, bucket=1 << 16):
        from .optim import FusedAdam
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs radnerf_b200.optim.FusedAdam (its step is capturable: no host reads)")
        self.model, self.opt, self.scaler, self.lambda_amb, self.phase = model, optimizer, scaler, lambda_amb, phase
        self.sync = sync if (sync is not None and sync.world > 1) else None
        if self.sync is not None:
            self.sync.overlap = False      # the backward is replayed from a graph: every reduction is issued after it
        self.bucket = int(bucket)
        self.graph = None
        self.graph_tail = None
        self.key = None
        self.static = None
        self.loss = None
        self.warm = False
        self.fallback_reason = None
        self.captures = 0
        self.replays = 0
        self.capture_ms = []       # host time of every (re-)capture, synchronisation and instantiation included
        self.pool = None           # every capture allocates from the same private pool: no cudaMalloc after the first one
        self.budget = None         # int32 [1] on the device: the marcher's sample budget (= padded mean_count)
        self.budget_value = None
        self.mc_version = None     # model.mean_count_version the budget was last refreshed for
        self.mc_host = 0           # last mean_count known to the HOST (decides the capacity bucket)
        self.mc_pending = None     # (pinned int32, event): an asynchronous read-back of a device-side mean_count
        self.capacity = None       # sample slots of the captured step's buffers
        self.max_samples = None    # rays x max_steps of the current batch shape: the marcher's hard bound (top rung of _capacity)
        self._params = [p for g in optimizer.param_groups for p in g["params"]]

    def _eager(self, batch):
        self.opt.device_groups = False
        return train_step(self.model, batch, self.opt, self.scaler, self.sync, self.lambda_amb, self.phase)

    def _load(self, batch):
        dev = self.static["rays_o"].device
        for k in self.KEYS:
            if self.static.get(k) is not None:
                self.static[k].copy_(torch.as_tensor(batch[k]), non_blocking=True)
        idx = batch.get("index", 0)
        self.static["index"].copy_(torch.as_tensor(idx, dtype=torch.long).reshape(-1)[:1].to(dev, non_blocking=True))

    def _capacity(self, mean_count=None):
        """(padded mean_count = the marcher's budget, capacity of the captured buffers).  The capacity has HEADROOM and HYSTERESIS: 20 %
        above the estimate at capture time, rounded up to `bucket` slots, and kept while the estimate stays inside [60 %, 100 %] of it --
        the estimate drifts by a few per cent at every occupancy update, and a re-capture (the step run eagerly under capture +
        instantiation: 10-100 ms depending on the host) must not happen every few updates."""
        padded = int(self.mc_host if mean_count is None else mean_count)
        padded = padded + (128 - padded % 128)            # raymarching._padded(mean_count, 128), what the eager marcher allocates
        cap = self.capacity
        if cap is None or padded > cap or padded < 0.6 * cap:
            cap = (int(padded * 1.2) + self.bucket - 1) // self.bucket * self.bucket
            # TOP RUNG: the marcher cannot emit more than rays x max_steps samples.  A capacity that comes within 25 % of that bound is
            # raised to the bound itself: the estimate keeps growing for the first occupancy updates of a run, and one more re-capture
            # (up to ~170 ms on a slow host, seen inside a timed window of the bench) costs more than the padding, which the kernels
            # skip (m_valid) -- from here on the step is never captured again for growth
            top = getattr(self, "max_samples", None)
            if top:
                top = (int(top) + 128 + self.bucket - 1) // self.bucket * self.bucket
                if cap >= 0.75 * top:
                    cap = top
        return padded, cap

    def _set_budget(self, padded):
        if self.budget is None:
            self.budget = torch.zeros(1, dtype=torch.int32, device=self.model.step_counter.device)
        if padded != self.budget_value:
            self.budget.fill_(padded)
            self.budget_value = padded

    def _follow_mean_count(self):
        """keep the device-side budget equal to the padded mean_count without stalling the host: a mean_count that update_extra_state
        left on the device is padded there and copied into `budget` in stream order; its value reaches the host through an asynchronous
        copy that is only LOOKED at once it has completed (it decides whether the capacity bucket must grow -- until then the marcher
        clamps to the current capacity, i.e. at worst drops a few more rays for a few steps)"""
        m = self.model
        v = getattr(m, "mean_count_version", 0)
        if v != self.mc_version:
            self.mc_version = v
            t = getattr(m, "_mean_count_dev", None)
            if t is None or self.budget is None or self.graph is None:
                self.mc_host = int(m.mean_count)           # host value (or the very first capture): plain path
                self.mc_pending = None
                self._set_budget(self._capacity()[0])
            else:
                self.budget.copy_(t + (128 - t % 128))
                self.budget_value = None
                host = torch.empty(1, dtype=torch.int32).pin_memory()
                host.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                self.mc_pending = (host, ev)
        if self.mc_pending is not None and self.mc_pending[1].query():
            self.mc_host = int(self.mc_pending[0].item())
            self.budget_value = self._capacity()[0]
            self.mc_pending = None

    def _capture(self, batch, capacity):
        import time
        from raymarching.raymarching import sample_budget
        t0 = time.perf_counter()
        m = self.model
        dev = batch["rays_o"].device
        previous = (self.graph, self.graph_tail)  # stay alive until the new graphs exist, so the shared pool keeps its memory
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        self.static = {k: (torch.as_tensor(batch[k]).to(dev).clone() if batch.get(k) is not None else None) for k in self.KEYS}
        self.static["index"] = torch.zeros(1, dtype=torch.long, device=dev)
        self._load(batch)
        if self.sync is not None and self.sync.flat is None:
            self.sync.flatten()
        if self.sync is None:
            # one GPU: the captured backward ASSIGNS its gradients (no pre-existing .grad -> autograd's AccumulateGrad keeps the incoming
            # tensor instead of adding into a zeroed one): 46 add launches and two 7 MB table additions per step gone.  The tensors live
            # in the graph's pool at fixed addresses; the optimiser's descriptor table is rebuilt for them inside the capture.
            for p in self._params:
                p.grad = None
        else:
            self.opt.prepare()             # data-parallel: gradients stay the in-place views of the exchange buffers
        self.opt.publish_groups()
        self.opt.device_groups = True
        if getattr(m, "_head_trainer", None) is not None:
            m._head_trainer.ensure(capacity)          # grow the fused step's buffers outside the capture
            m._head_trainer.packed_for = None         # the captured step must contain the table / weight re-packing launches
        local_step = m.local_step
        self.counter_row = local_step % 16
        torch.cuda.synchronize(dev)
        _ = m.mean_count                   # run_cuda reads it as a Python int: make the host copy NOW, a capture cannot synchronise
        g, g_tail = torch.cuda.CUDAGraph(), None
        try:
            with sample_budget(capacity, self.budget):
                if self.sync is None:
                    with torch.cuda.graph(g, pool=self.pool):
                        self.loss = train_step(m, self.static, self.opt, self.scaler, None, self.lambda_amb, self.phase)
                else:
                    with torch.cuda.graph(g, pool=self.pool):
                        self.loss = forward_backward(m, self.static, self.opt, self.scaler, self.lambda_amb, self.phase).detach()
                    g_tail = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_tail, pool=self.pool):
                        optimizer_tail(m, self.static, self.opt, self.scaler)
        finally:
            m.local_step = local_step      # capture ran the host bookkeeping of one step without executing it
        self.graph, self.graph_tail = g, g_tail
        del previous
        self.captures += 1
        self.capture_ms.append((time.perf_counter() - t0) * 1e3)

    def __call__(self, batch):
        m = self.model
        amp = bool(m.opt.fp16)
        shapes = tuple((k, tuple(torch.as_tensor(batch[k]).shape)) for k in self.KEYS if batch.get(k) is not None)
        if self.fallback_reason is not None or not amp:
            return self._eager(batch)
        if not self.warm:        # (the only place the host reads mean_count itself: the cold regime synchronises every step anyway)
            self.warm = m.mean_count > 0                   # one eager step in the steady regime: lazy initialisations, grads exist
            return self._eager(batch)
        self._follow_mean_count()
        self.max_samples = int(torch.as_tensor(batch["rays_o"]).numel() // 3) * int(m.opt.max_steps)
        padded, capacity = self._capacity()
        key = (capacity, shapes, self.phase)
        if key != self.key:
            if os.environ.get("RADNERF_DEBUG_CAPTURE"):
                print("GraphedTrainStep capture: mean_count(host) %s padded %d capacity %s -> %d" % (self.mc_host, padded, self.capacity, capacity), flush=True)
            try:
                self._capture(batch, capacity)
                self.key, self.capacity = key, capacity
            except Exception as e:  # noqa: BLE001
                self.fallback_reason = repr(e)[:300]
                self.graph, self.graph_tail, self.key = None, None, None
                torch.cuda.synchronize()
                return self._eager(batch)
        else:
            self._load(batch)
            self.opt.publish_groups()
        if self.sync is not None:
            self.sync.begin_step()
        self.graph.replay()
        if self.graph_tail is not None:
            self.sync.finish()
            self.graph_tail.replay()
        self.replays += 1
        # the replay rewrote every parameter through frozen pointers: consumers that cache derived data per (pointer, version) --
        # the fp16 tables / weight blobs of the fused frame and of the fused density query -- must see the step
        torch.autograd.graph.increment_version(self._params)
        # host bookkeeping of run_cuda's training branch: the counter row of this step, the step index
        row = m.local_step % 16
        if row != self.counter_row:
            m.step_counter[row].copy_(m.step_counter[self.counter_row])
        m.local_step += 1
        return self.loss.detach()
