"""Streaming audio front-end hand-off: ASR feature rows in, the renderer's [8, dim, 16] attention
window out, without leaving the device.

The reference keeps the ASR logits in a ring `feat_queue [slots * context, dim]` that the ASR thread fills slot by slot
(nerf/asr.py:217-224) and, per video frame, slides a 16-row window forward by 2 rows, keeping the 8 latest windows
(`get_next_feat`, nerf/asr.py:160-183, started with front = size - 8, tail = 8 and four all-zero windows, :103-109).
Here the ring lives on the device and a frame's 8 windows are ONE gather (8 x 16 row indices modulo the ring size) written
straight into a destination of the caller's choice -- e.g. the audio slice of a frame lane's input block
(radnerf_b200.stream: flat[24:]) -- instead of slices, a permute per window, a Python list and a stack per frame."""
import torch


class FeatureRing:
    WINDOW, HOP, DEPTH = 16, 2, 8

    def __init__(self, slots, context, dim, device):
        self.context, self.dim, self.slots = context, dim, slots
        self.size = slots * context
        self.queue = torch.zeros(self.size, dim, dtype=torch.float32, device=device)
        self.slot = 0
        self.front = self.size - self.WINDOW // 2       # the reference's start: the first window is rows [size - 8, 8)
        self.calls = 0
        # the 8 latest windows, oldest first, as (first ring row | None for the start-up zero windows, snapshot | None).
        # In the reference a window that does not wrap around the ring end is a VIEW of the ring (a slice + permute kept in a
        # list, nerf/asr.py:165-176): it shows later overwrites when it is finally stacked; a wrapping window is a torch.cat,
        # i.e. a snapshot.  Both behaviours are kept.
        self.windows = [(None, None)] * self.DEPTH
        self._offsets = torch.arange(self.WINDOW, device=device)

    def push(self, feats):
        """one ASR context worth of rows [n <= context, dim] into the next slot (nerf/asr.py:221-224)"""
        start = self.slot * self.context
        self.queue[start:start + feats.shape[0]] = feats.to(self.queue.device, torch.float32)
        self.slot = (self.slot + 1) % self.slots

    def next_window(self, out=None):
        """the [8, dim, 16] block of the next video frame (== ASR.get_next_feat()), optionally written into `out` (any tensor
        with 8*dim*16 elements, e.g. the audio slice of a frame lane's input block)"""
        for _ in range(self.DEPTH // 2 if self.calls == 0 else 1):     # the first frame brings four windows, later frames one
            start = self.front
            snapshot = None
            if start + self.WINDOW > self.size:                        # wraps: the reference concatenates -> a copy
                snapshot = self.queue[(start + self._offsets) % self.size].clone()
            self.windows = self.windows[1:] + [(start, snapshot)]
            self.front = (self.front + self.HOP) % self.size
        self.calls += 1
        starts = torch.tensor([0 if w[0] is None else w[0] for w in self.windows], device=self.queue.device)
        block = self.queue[(starts[:, None] + self._offsets[None, :]) % self.size]     # one gather: [8, 16, dim], live rows
        for i, (start, snapshot) in enumerate(self.windows):
            if start is None:
                block[i].zero_()
            elif snapshot is not None:
                block[i].copy_(snapshot)
        block = block.permute(0, 2, 1)                                                 # [8, dim, 16]
        if out is None:
            return block.contiguous()
        out.view(self.DEPTH, self.dim, self.WINDOW).copy_(block)
        return out
