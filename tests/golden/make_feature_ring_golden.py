"""Golden vectors for SURVEY 8(f) rank 4, produced by the REFERENCE's own ASR.get_next_feat (nerf/asr.py:160-183).

nerf/asr.py is imported from /root/reference with its audio-IO packages stubbed (pyaudio, soundfile, resampy are not
installed; they are not touched by the method under test).  `get_next_feat` runs UNBOUND on a bare object that carries
exactly the attributes ASR.__init__ sets up for it (asr.py:100-109), on CPU tensors; ring writes are the statement of
run_step (asr.py:221-224).  The scenario comes from tests/feature_ring_case.py.

    python tests/golden/make_feature_ring_golden.py     ->  tests/golden/feature_ring.npz   (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from feature_ring_case import CASES, script   # noqa: E402


def main():
    import importlib.machinery
    from transformers import AutoModelForCTC, AutoProcessor   # noqa: F401  (resolve transformers' lazy imports before stubbing)
    for name in ("pyaudio", "soundfile", "resampy"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
                sys.modules[name].__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.path.append("/root/reference")
    from nerf.asr import ASR
    out = {}
    for case in CASES:
        slots, context, dim, ops = script(case)
        a = types.SimpleNamespace()
        a.feat_buffer_size, a.context_size, a.audio_dim, a.feat_buffer_idx = slots, context, dim, 0
        a.feat_queue = torch.zeros(slots * context, dim, dtype=torch.float32)
        a.front = slots * context - 8
        a.tail = 8
        a.att_feats = [torch.zeros(dim, 16, dtype=torch.float32)] * 4
        blocks = []
        for op in ops:
            if op[0] == "write":
                feats = torch.from_numpy(op[1])
                start = a.feat_buffer_idx * a.context_size
                a.feat_queue[start:start + feats.shape[0]] = feats
                a.feat_buffer_idx = (a.feat_buffer_idx + 1) % a.feat_buffer_size
            else:
                blocks.append(ASR.get_next_feat(a).numpy().copy())
        out[case] = np.stack(blocks)
    # the blocks are highly redundant (each window appears in 8 consecutive frames): store every frame's NEWEST window plus
    # the complete first 12 and last 4 blocks, and a checksum of every block
    packed = {}
    for case, b in out.items():
        packed[case + "/newest"] = b[:, -1]
        packed[case + "/head"] = b[:12]
        packed[case + "/tail"] = b[-4:]
        packed[case + "/sums"] = b.astype(np.float64).sum(axis=(2, 3))
    np.savez_compressed(os.path.join(HERE, "feature_ring.npz"), **packed)
    print({k: v.shape for k, v in packed.items()})


if __name__ == "__main__":
    main()
