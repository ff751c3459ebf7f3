"""Build the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

The reference (Karthik-Ragunath/RAD-NeRF) ships its hot path as four torch CUDA
extensions and nothing else -- there is no CPU implementation.  This recipe compiles
those sources *where they lie* under /root/reference (nothing is copied into the repo)
with the reference's own flags (gridencoder/backend.py:6-9, freqencoder/backend.py:9),
except `-std=c++14` -> `-std=c++17` (torch 2.11 headers need it) and an explicit
sm_100a gencode.  Outputs (.so + ninja scratch) go to oracle/_ref/ only; that
directory is git-ignored but travels to the GPU box with the snapshot.

The resulting modules are used by:
  * oracle/make_golden.py  -- runs the reference kernels on a B200 to produce the
                              golden vectors committed under tests/golden/
  * tests/ (-m gpu)        -- optional live cross-check when the .so files are present
  * bench.py               -- the `ref_cuda` comparison arm (reference kernels timed
                              on the same GPU)
Never imported by the product path.
"""
import os
import sys
import shutil

REF = os.environ.get("RADNERF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    # module name (what the reference's wrappers `import`), sources, extra nvcc flags
    "_gridencoder": (["gridencoder/src/gridencoder.cu", "gridencoder/src/bindings.cpp"], []),
    "_raymarching_face": (["raymarching/src/raymarching.cu", "raymarching/src/bindings.cpp"], []),
    "_freqencoder": (["freqencoder/src/freqencoder.cu", "freqencoder/src/bindings.cpp"], ["-use_fast_math"]),
    "_shencoder": (["shencoder/src/shencoder.cu", "shencoder/src/bindings.cpp"], []),
}


def build_one(name):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    srcs, extra = EXTS[name]
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    nvcc_flags = ["-O3", "-std=c++17",
                  "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                  "-U__CUDA_NO_HALF2_OPERATORS__",
                  "-gencode", "arch=compute_100a,code=sm_100a"] + extra
    load(name=name, extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=nvcc_flags,
         sources=[os.path.join(REF, s) for s in srcs], build_directory=bdir,
         verbose=True, is_python_module=False)
    so = os.path.join(bdir, name + ".so")
    shutil.copy2(so, os.path.join(OUT, name + ".so"))
    print("built", os.path.join(OUT, name + ".so"))


def main():
    if not os.path.isdir(REF):
        print("reference tree not present at", REF, "- nothing to build")
        return 0
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(EXTS)
    for n in names:
        if os.path.exists(os.path.join(OUT, n + ".so")) and not os.environ.get("FORCE"):
            print("up to date:", n)
            continue
        build_one(n)
    return 0


if __name__ == "__main__":
    sys.exit(main())
