"""ctypes binding of libradnerf_b200.so (C ABI: include/radnerf_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load, importing any operator
package raises.  PyTorch is used only for device memory and streams; tensors cross the boundary as raw pointers.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libradnerf_b200.so")

_u32, _f32, _vp, _i = C.c_uint32, C.c_float, C.c_void_p, C.c_int

# name -> argtypes, exactly the prototypes of include/radnerf_b200.h
_PROTOS = {
    "rn_grid_encode_forward": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _vp, _u32, _u32, _u32, _u32, _u32, _vp],
    "rn_grid_encode_backward": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp],
    "rn_grad_total_variation": [_vp, _vp, _vp, _vp, _f32, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _u32, _u32, _vp],
    "rn_grid_level_geometry": [_f32, _u32, _u32, _vp, _vp, _vp],
    "rn_near_far_from_aabb": [_vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp],
    "rn_sph_from_ray": [_vp, _vp, _f32, _u32, _vp, _vp],
    "rn_morton3D": [_vp, _u32, _vp, _vp],
    "rn_morton3D_invert": [_vp, _u32, _vp, _vp],
    "rn_packbits": [_vp, _u32, _f32, _vp, _vp],
    "rn_morton3D_dilation": [_vp, _u32, _u32, _vp, _vp],
    "rn_march_rays_train": [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rn_march_rays_train_budget": [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rn_march_rays_train_backward": [_vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp],
    "rn_composite_rays_train_forward": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp],
    "rn_composite_rays_train_backward": [_vp] * 11 + [_u32, _u32, _f32, _vp, _vp, _vp, _vp],
    "rn_march_rays": [_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rn_composite_rays": [_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rn_freq_encode_forward": [_vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "rn_freq_encode_backward": [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "rn_sh_encode_forward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp],
    "rn_sh_encode_backward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp],
    "rn_get_rays": [_vp, _f32, _f32, _f32, _f32, _u32, _u32, _vp, _u32, _vp, _vp, _vp],
    "rn_scatter_rows_to_peers": [_vp, _vp, _u32, _u32, _vp, _u32, _vp],
    "rn_image_to_uint8": [_vp, _vp, C.c_uint64, _vp],
    "rn_occupancy_pack": [_vp, _u32, _u32, _f32, _vp, _vp],
    "rn_packbits_min": [_vp, _u32, _f32, _vp, _vp, _vp],
    "rn_occupancy_merge": [_vp, _vp, _u32, _f32, _vp, _vp, _vp],
    "rn_scatter_rows_to_root": [_vp, _vp, _u32, _u32, _vp, _vp, _u32, _u32, _u32, _u32, C.c_uint64, _vp],
    "rn_stage_frame_at_root": [_vp, _vp, C.c_uint64, _u32, _vp, _u32, _u32, _u32, C.c_uint64, _u32, _vp, _vp],
    "rn_selftest_umma": [_vp, _vp, _vp, _u32, _u32, _vp],
    "rn_selftest_umma_mn": [_vp, _vp, _vp, _u32, _u32, _u32, _vp],
    "rn_grid_backward3": [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp],
}

_lib = None


def lib():
    """Load (once) and return the CUDA library.  Raises if it is missing -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python rad-nerf_b200/build.py` (or __graft_entry__.build()). "
                "radnerf_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in _PROTOS.items():
            fn = getattr(L, name)  # AttributeError here == ABI drift; fail loudly
            fn.argtypes = args
            fn.restype = _i
        L.rn_last_error_string.restype = C.c_char_p
        L.rn_abi_version.restype = _i
        L.rn_launch_count.restype = C.c_uint64
        for name in _EXTRA_PROTOS:
            fn = getattr(L, name)
            fn.argtypes, fn.restype = _EXTRA_PROTOS[name]
        _lib = L
    return _lib


# prototypes registered by other modules (frame renderer) before first load
_EXTRA_PROTOS = {}


def register(name, argtypes, restype=_i):
    _EXTRA_PROTOS[name] = (argtypes, restype)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.argtypes, fn.restype = argtypes, restype


def check(rc, what=""):
    if rc != 0:
        msg = lib().rn_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(msg or f"{what} failed with code {rc}")


def cur_stream():
    """cudaStream_t of torch's current stream on the current device, as the ABI's `void* stream`."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (honours storage offsets); None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def call(name, *args):
    """Invoke ABI function `name` on torch's current stream: tensors become device pointers (None -> NULL), scalars pass
    through, the stream is appended, a non-zero return code raises RuntimeError with the library's message."""
    fn = getattr(lib(), name)
    dev = None
    for a in args:   # every tensor argument lives on ONE device; the launch goes to that device's current stream
        if torch.is_tensor(a):
            if not a.is_cuda:
                raise RuntimeError("radnerf_b200.%s: expected CUDA tensors (this library has no CPU path)" % name)
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise RuntimeError("radnerf_b200.%s: tensor arguments on different devices (%s, %s)" % (name, dev, a.device))
    argv = [ptr(a) if (a is None or torch.is_tensor(a)) else a for a in args]
    if dev is None or dev.index == torch.cuda.current_device():
        check(fn(*argv, cur_stream()), name)
    else:
        with torch.cuda.device(dev):
            check(fn(*argv, cur_stream()), name)


def rows(t, width=None, dtype=torch.float32):
    """[..., width] tensor -> (contiguous 2-D `dtype` view on the device, leading shape); the library has no CPU path"""
    require_cuda(t)
    lead = tuple(t.shape[:-1])
    flat = t.reshape(-1, t.shape[-1] if width is None else width)
    if dtype is not None and flat.dtype != dtype:
        flat = flat.to(dtype)
    return flat.contiguous(), lead


def launch_count():
    return int(lib().rn_launch_count())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("radnerf_b200: expected a CUDA tensor (this library has no CPU path)")
