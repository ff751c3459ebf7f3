"""The seeded streaming scenario shared by the feature-ring golden generator and its tests: which calls happen in which
order (ring writes by the ASR side, window reads by the render side) and what is written."""
import numpy as np

CASES = {
    # name: (slots, context, dim, frames).  The GUI runs two ASR steps per video frame (nerf/gui.py:560-564); a context of m
    # audio frames is written once every m steps, so a write lands every m/2 video frames.
    "wav2vec_m50": (4, 50, 44, 260),       # the reference defaults: feat_buffer_size 4, -m 50, esperanto wav2vec (44-d)
    "deepspeech_m10": (4, 10, 29, 120),    # 40-row ring: every 20th window wraps, writes overtake live windows
    "short_last": (4, 12, 32, 90),         # a final context shorter than m rows
}


def script(name):
    """-> (slots, context, dim, [("write", feats) | ("read",)])"""
    slots, context, dim, frames = CASES[name]
    rng = np.random.default_rng(sum(name.encode()))
    ops = []
    steps = 0
    for f in range(frames):
        for _ in range(2):
            steps += 1
            if steps % context == 0:
                n = context - 3 if (name == "short_last" and steps // context == 5) else context
                ops.append(("write", rng.standard_normal((n, dim)).astype(np.float32)))
        ops.append(("read",))
    return slots, context, dim, ops
