"""Load the reference's OWN, unmodified classes (baseline/_ref/reference/nerf/network.py, nerf/renderer.py, nerf/utils.py --
installed by baseline/install_ref.py) on top of either operator stack:

    backend="ours"  `raymarching`, `gridencoder`, `freqencoder`, `shencoder`, `encoding`, `activation` resolve to the drop-in
                    packages of this repository (rad-nerf_b200/) -- the north_star claim "nerf/network.py and nerf/renderer.py
                    run unchanged on it";
    backend="ref"   they resolve to the reference's own wrappers over the reference's own compiled CUDA extensions
                    (oracle/_ref/_gridencoder.so, _raymarching_face.so, _freqencoder.so, _shencoder.so, built from
                    /root/reference by oracle/build_ref.py) -- the yardstick.

Both stacks can live in one process: every load imports a private copy of the module set and removes it from sys.modules
again, so `load("ours").NeRFNetwork` and `load("ref").NeRFNetwork` are two distinct classes from the same source file whose
globals point at different operator packages.

TEST / BENCH INFRASTRUCTURE: used by tests/ and by `bench.py --impl reference`; the product never imports it.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_PY = os.path.join(HERE, "_ref", "reference")
OURS = os.path.join(ROOT, "rad-nerf_b200")
REF_SO = os.path.join(ROOT, "oracle", "_ref")

_OP_PACKAGES = ("raymarching", "gridencoder", "freqencoder", "shencoder", "encoding", "activation")
_PRIVATE = _OP_PACKAGES + ("nerf", "_gridencoder", "_raymarching_face", "_freqencoder", "_shencoder")
# pip packages the reference imports at module level but never touches on the hot path; absent offline
_STUBS = ("trimesh", "tensorboardX", "matplotlib", "matplotlib.pyplot", "mcubes", "imageio", "lpips", "pyaudio", "soundfile",
          "resampy", "dearpygui", "dearpygui.dearpygui", "face_alignment", "python_speech_features", "configargparse")
_cache = {}


def available(backend="ours"):
    if not os.path.exists(os.path.join(REF_PY, "nerf", "network.py")):
        return False
    if backend == "ref":
        return all(os.path.exists(os.path.join(REF_SO, n + ".so")) for n in ("_gridencoder", "_raymarching_face", "_freqencoder", "_shencoder"))
    return True


def _stub_missing():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:   # noqa: BLE001  (absent or broken offline: an empty module satisfies the import statement)
            m = types.ModuleType(name)
            if name == "tensorboardX":
                m.SummaryWriter = object
            sys.modules[name] = m
    if "torch_ema" not in sys.modules:
        try:
            importlib.import_module("torch_ema")
        except Exception:   # noqa: BLE001
            m = types.ModuleType("torch_ema")

            class ExponentialMovingAverage:   # only ever constructed by Trainer, which the hot path does not use
                def __init__(self, *a, **k):
                    raise RuntimeError("torch_ema is not installed")
            m.ExponentialMovingAverage = ExponentialMovingAverage
            sys.modules["torch_ema"] = m


def load(backend="ours"):
    """-> namespace(NeRFNetwork, network, renderer, utils, raymarching, encoding, gridencoder, ...) of the stock classes bound to `backend`"""
    if backend in _cache:
        return _cache[backend]
    if backend not in ("ours", "ref"):
        raise ValueError(backend)
    if not available(backend):
        raise RuntimeError("stock reference not installed (run baseline/install_ref.py; backend='ref' also needs oracle/build_ref.py)")
    _stub_missing()
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in _PRIVATE}
    saved_path = list(sys.path)
    try:
        # the reference tree must not shadow anything else (it has a top-level test.py / main.py): it goes last, the operator
        # packages of the chosen stack first
        sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (OURS, REF_PY, REF_SO, "/root/reference")]
        if backend == "ours":
            sys.path[:0] = [OURS]
            sys.path.append(REF_PY)
        else:
            sys.path[:0] = [REF_SO]           # _gridencoder.so ... : `import _gridencoder as _backend` in the wrappers succeeds
            sys.path.insert(1, REF_PY)        # the reference's own wrapper packages
        ns = types.SimpleNamespace(backend=backend)
        for name in _OP_PACKAGES:
            setattr(ns, name, importlib.import_module(name))
        ns.network = importlib.import_module("nerf.network")
        ns.renderer = importlib.import_module("nerf.renderer")
        ns.utils = importlib.import_module("nerf.utils")
        ns.NeRFNetwork = ns.network.NeRFNetwork
        origin = os.path.dirname(os.path.abspath(ns.raymarching.__file__))
        expect = os.path.join(OURS if backend == "ours" else REF_PY, "raymarching")
        if origin != expect:
            raise RuntimeError("operator packages resolved to %s, expected %s" % (origin, expect))
        if os.path.dirname(os.path.abspath(ns.network.__file__)) != os.path.join(REF_PY, "nerf"):
            raise RuntimeError("nerf.network did not come from the installed reference copy")
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _PRIVATE]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
    _cache[backend] = ns
    return ns


def default_opt(**over):
    """the argparse namespace of `test.py -O --torso` restricted to what NeRFRenderer / NeRFNetwork read (main.py:12-120)"""
    opt = dict(bound=1.0, min_near=0.05, density_thresh=10.0, density_thresh_torso=0.01, dt_gamma=1 / 256, max_steps=16,
               exp_eye=True, fp16=True, torso=True, smooth_lips=True, test_train=False, cuda_ray=True, train_camera=False, att=2, emb=False,
               ind_num=10000, ind_dim=4, ind_dim_torso=8, amb_dim=2, torso_shrink=0.8,
               asr_model="cpierse/wav2vec2-large-xlsr-53-esperanto", num_rays=4096 * 16, update_extra_interval=16,
               finetune_lips=False, patch_size=1, lambda_amb=0.1, color_space="srgb", bg_img="", fbg=False, fix_eye=-1, part=False, part2=False)
    opt.update(over)
    return types.SimpleNamespace(**opt)


def build(backend, device, opt=None, seed=0, **over):
    """a stock NeRFNetwork on `device`, seeded so that both backends get identical parameters"""
    import torch
    st = load(backend)
    torch.manual_seed(seed)
    net = st.NeRFNetwork(opt or default_opt(**over))
    return net.to(device)
