"""Streaming audio front-end hand-off: ASR feature rows in, the renderer's [8, dim, 16] attention
window out, without leaving the device.

The reference keeps the ASR logits in a ring `feat_queue [slots * context, dim]` that the ASR thread fills slot by slot
(nerf/asr.py:217-224) and, per video frame, slides a 16-row window forward by 2 rows, keeping the 8 latest windows
(`get_next_feat`, nerf/asr.py:160-183, started with front = size - 8, tail = 8 and four all-zero windows, :103-109).
Here the ring lives on the device and a frame's 8 windows are ONE kernel launch (`rn_feature_window`,
csrc/feature_ring.cu) written straight into a destination of the caller's choice -- e.g. the audio slice of a frame lane's
input block (radnerf_b200.stream: flat[24:]) -- instead of 8 slices, 8 permutes, a Python list, a stack and a copy per frame.

`WindowBook` is the host-side bookkeeping (pure Python, no tensors); `FeatureRing` owns the device buffers.  There is no CPU
path: the ring must live on a CUDA device."""
import ctypes as C

import torch

from . import abi

WINDOW, HOP, DEPTH = 16, 2, 8


class RingWindows(C.Structure):     # rn_ring_windows
    _fields_ = [("start", C.c_int32 * DEPTH), ("snapshot", C.c_int32 * DEPTH), ("fresh", C.c_int32 * DEPTH)]


abi.register("rn_feature_window", [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(RingWindows), C.c_void_p, C.c_void_p, C.c_void_p])


class WindowBook:
    """Which 8 windows the next video frame sees.  In the reference a window that does not wrap around the ring end is a
    VIEW of the ring (a slice + permute kept in a list, nerf/asr.py:165-176): it shows later overwrites of the ring when it
    is finally stacked; a wrapping window is a torch.cat, i.e. a snapshot taken when the window was created.  Both
    behaviours are kept: a window is (start row | -1 for the start-up zero windows, snapshot slot | -1 for live)."""

    def __init__(self, size):
        if size < WINDOW:
            raise ValueError("the ring must hold at least one 16-row window")
        self.size = size
        self.front = size - WINDOW // 2          # the reference's start: the first window is rows [size - 8, 8)
        self.windows = [(-1, -1)] * (DEPTH // 2)  # four zero windows (asr.py:109)

    def advance(self):
        """-> (start[8], snapshot[8], fresh[8]) for this frame, oldest window first; moves the book one frame on"""
        fresh = []
        while len(self.windows) < DEPTH:                       # the first frame brings four windows, later frames one
            start = self.front
            slot = -1
            if start + WINDOW >= self.size:                    # `front < tail` fails (asr.py:166): torch.cat -> a copy
                used = {s for _, s in self.windows if s >= 0}
                slot = min(set(range(DEPTH)) - used)
            self.windows.append((start, slot))
            fresh.append(len(self.windows) - 1)
            self.front = (self.front + HOP) % self.size
        out = ([w[0] for w in self.windows], [w[1] for w in self.windows],
               [1 if (k in fresh and self.windows[k][1] >= 0) else 0 for k in range(DEPTH)])
        self.windows = self.windows[1:]                        # discard the oldest (asr.py:181)
        return out


class FeatureRing:
    WINDOW, HOP, DEPTH = WINDOW, HOP, DEPTH

    def __init__(self, slots, context, dim, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("radnerf_b200: FeatureRing needs a CUDA device (this library has no CPU path)")
        self.context, self.dim, self.slots = context, dim, slots
        self.size = slots * context
        self.queue = torch.zeros(self.size, dim, dtype=torch.float32, device=device)          # feat_queue (asr.py:103)
        self.snapshots = torch.zeros(DEPTH, WINDOW, dim, dtype=torch.float32, device=device)
        self.slot = 0
        self.book = WindowBook(self.size)

    def push(self, feats):
        """one ASR context worth of rows [n <= context, dim] into the next slot (nerf/asr.py:221-224)"""
        start = self.slot * self.context
        self.queue[start:start + feats.shape[0]].copy_(feats, non_blocking=True)
        self.slot = (self.slot + 1) % self.slots

    def next_window(self, out=None):
        """the [8, dim, 16] block of the next video frame (== ASR.get_next_feat()), optionally written into `out` (any
        contiguous fp32 device tensor with 8*dim*16 elements, e.g. the audio slice of a frame lane's input block)"""
        if out is None:
            out = torch.empty(DEPTH, self.dim, WINDOW, dtype=torch.float32, device=self.queue.device)
        elif out.numel() != DEPTH * self.dim * WINDOW or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous fp32 tensor of 8 * dim * 16 elements")
        abi.require_cuda(out)
        start, snapshot, fresh = self.book.advance()
        w = RingWindows()
        w.start[:], w.snapshot[:], w.fresh[:] = start, snapshot, fresh
        with torch.cuda.device(self.queue.device):
            abi.check(abi.lib().rn_feature_window(abi.ptr(self.queue), self.size, self.dim, C.byref(w), abi.ptr(self.snapshots),
                                                  abi.ptr(out), abi.cur_stream()), "rn_feature_window")
        return out
