"""radnerf_b200 -- host side of the B200-native RAD-NeRF hot path (ctypes over libradnerf_b200.so).

The sibling top-level packages (`gridencoder`, `raymarching`, `freqencoder`, `shencoder`, `encoding`, `activation`)
mirror the reference's Python operator API; this package holds the library binding (`abi`), synthetic input
generators (`synthetic`) and the fused frame renderer.
"""
import os as _os

# Frame lanes (radnerf_b200.stream) keep 4-8 frames in flight on separate streams, each with a conditioning / copy / torso side branch:
# more independent queues than the driver's default 8 hardware connections.  Streams that share a connection serialise ("false
# dependencies"): measured on one B200 with a 184x184-ray frame (one rank's share of a 512x512 frame on 8 GPUs), 8 lanes deliver
# 8.6 k frames/s with 8 connections and 17.9 k with 32 (tools/lane_probe.py, profiles/r02_lane_probe.json).  The variable is read when
# the CUDA context is created, so it is set at import -- before the first CUDA call of a process that imports this package first;
# an explicit setting of the user's is respected.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import abi  # noqa: F401,E402
from .abi import check, cur_stream, launch_count, LIB_PATH  # noqa: F401


def lib():
    return abi.lib()
