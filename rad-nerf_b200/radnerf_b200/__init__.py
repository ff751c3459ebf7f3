"""radnerf_b200 -- host side of the B200-native RAD-NeRF hot path (ctypes over libradnerf_b200.so).

The sibling top-level packages (`gridencoder`, `raymarching`, `freqencoder`, `shencoder`, `encoding`, `activation`)
mirror the reference's Python operator API; this package holds the library binding (`abi`), synthetic input
generators (`synthetic`) and the fused frame renderer.
"""
from . import abi  # noqa: F401
from .abi import check, cur_stream, launch_count, LIB_PATH  # noqa: F401


def lib():
    return abi.lib()
