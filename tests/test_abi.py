"""The C-ABI library loads and exports every symbol include/radnerf_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "radnerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    # one entry point per reference binding (gridencoder 3, raymarching 12, freqencoder 2, shencoder 2 = 19)
    for s in ["rn_grid_encode_forward", "rn_grid_encode_backward", "rn_grad_total_variation", "rn_near_far_from_aabb",
              "rn_sph_from_ray", "rn_morton3D", "rn_morton3D_invert", "rn_packbits", "rn_morton3D_dilation",
              "rn_march_rays_train", "rn_march_rays_train_budget", "rn_march_rays_train_backward", "rn_composite_rays_train_forward",
              "rn_composite_rays_train_backward", "rn_march_rays", "rn_composite_rays", "rn_freq_encode_forward",
              "rn_freq_encode_backward", "rn_sh_encode_forward", "rn_sh_encode_backward"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from radnerf_b200 import abi as L
    assert os.path.exists(L.LIB_PATH)
    cdll = ctypes.CDLL(L.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(cdll, s)]
    assert not missing, f"declared in include/radnerf_b200.h but not exported: {missing}"
    cdll.rn_abi_version.restype = ctypes.c_int
    assert cdll.rn_abi_version() >= 1
    # the Python binding table matches the header
    assert set(L._PROTOS) <= set(declared_symbols())


def test_argument_errors_are_reported_without_a_gpu():
    from radnerf_b200 import abi as L
    lib = L.lib()
    # unsupported C -> RN_E_UNSUPPORTED with a message naming the constraint (reference: runtime_error "C must be 1, 2, 4, or 8")
    rc = lib.rn_grid_encode_forward(None, None, None, None, 0, 3, 2, 0, 0.5, 16, None, 0, 0, 0, 0, 1, None)
    assert rc == -1  # L == 0 is a bad argument
    rc = lib.rn_freq_encode_forward(None, 4, 2, 10, 41, None, None)
    assert rc == -1 and b"output_dim" in lib.rn_last_error_string()
    rc = lib.rn_sh_encode_forward(None, None, 4, 3, 9, None, None)
    assert rc == -2 and b"degree" in lib.rn_last_error_string()
    # the fused-frame / streaming entry points validate their descriptors before touching the device
    import ctypes as C
    from radnerf_b200 import frame, stream
    assert lib.rn_frame_head(None, None) == -1 and b"null descriptor" in lib.rn_last_error_string()
    hd = frame.FrameHeadDesc()
    hd.N = 64
    assert lib.rn_frame_head(C.byref(hd), None) == -1 and b"null pointer" in lib.rn_last_error_string()
    assert lib.rn_frame_conditioning(None, None) == -1
    s = stream.LaneSubmit()
    s.phase = 1
    assert lib.rn_lane_submit_frame(C.byref(s)) == -1 and b"phase 1 needs" in lib.rn_last_error_string()
    assert lib.rn_scatter_rows_to_peers(None, None, 8, 3, None, 2, None) == -1
    assert lib.rn_image_to_uint8(None, None, 0, None) == 0       # empty input is a no-op, as everywhere in the ABI


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rad-nerf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_ctypes_mirrors_match_the_library_struct_layouts():
    """the ctypes.Structure mirrors in frame.py / stream.py must have exactly the size of the C structs they stand for"""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
    from radnerf_b200 import abi, frame, stream
    L = abi.lib()
    L.rn_sizeof.restype = C.c_uint32
    L.rn_sizeof.argtypes = [C.c_char_p]
    for name, mirror in (("rn_grid_table", frame.GridTable), ("rn_conditioning_desc", frame.ConditioningDesc),
                         ("rn_frame_head_desc", frame.FrameHeadDesc), ("rn_frame_torso_desc", frame.FrameTorsoDesc),
                         ("rn_lane_submit", stream.LaneSubmit)):
        assert L.rn_sizeof(name.encode()) == C.sizeof(mirror) > 0, name
    assert L.rn_sizeof(b"no_such_struct") == 0
