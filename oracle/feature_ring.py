"""CPU restatement of the reference's ASR feature ring and window hand-off.  TEST INFRASTRUCTURE ONLY (see oracle.py).

Follows nerf/asr.py: the ring and its cursors (:100-109), the slot-by-slot write of run_step (:217-224), get_next_feat
(:160-183).  Parity status: pinned against tests/golden/feature_ring.npz, produced by the reference's own
ASR.get_next_feat (tests/golden/make_feature_ring_golden.py), checked by tests/test_feature_ring.py."""
import numpy as np


class Ring:
    def __init__(self, slots, context, dim):
        self.context, self.slots = context, slots
        self.queue = np.zeros((slots * context, dim), np.float32)      # feat_queue, asr.py:103
        self.idx = 0                                                   # feat_buffer_idx
        self.front = slots * context - 8                               # asr.py:106-107
        self.tail = 8
        # att_feats (asr.py:109): four zero windows.  A later entry is ("view", front, tail) -- torch's slice + permute is a
        # view of the ring and is only read when the list is stacked -- or ("copy", array) for torch.cat
        self.windows = [("copy", np.zeros((16, dim), np.float32))] * 4

    def write(self, feats):                                            # asr.py:221-224
        start = self.idx * self.context
        self.queue[start:start + feats.shape[0]] = feats
        self.idx = (self.idx + 1) % self.slots

    def next_window(self):                                             # asr.py:160-183
        n = self.queue.shape[0]
        while len(self.windows) < 8:
            if self.front < self.tail:
                w = ("view", self.front, self.tail)
            else:
                w = ("copy", np.concatenate([self.queue[self.front:], self.queue[:self.tail]], axis=0))
            self.front = (self.front + 2) % n
            self.tail = (self.tail + 2) % n
            self.windows.append(w)
        rows = [self.queue[w[1]:w[2]] if w[0] == "view" else w[1] for w in self.windows]
        out = np.stack([r.T for r in rows], axis=0)                    # [8, dim, 16]
        self.windows = self.windows[1:]
        return out
