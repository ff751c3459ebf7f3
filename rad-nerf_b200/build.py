"""Build libradnerf_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch headers are involved: each .cu compiles in seconds.  The shared object is written next to this file so
that it travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libradnerf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# per-file extras.  freqencoder mirrors the reference's -use_fast_math (freqencoder/backend.py:9) so __sinf/scalbnf
# expand to the same instructions; the marchers must NOT be built with fast math (bit-exact contract).
EXTRA = {"freqencoder.cu": ["-use_fast_math"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "radnerf_b200.h"),
                                                                 os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def compile_one(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + [os.path.join(CSRC, src),
                                                                                         os.path.join(HERE, "..", "include", "radnerf_b200.h")]
    if os.path.exists(obj) and all(os.path.getmtime(d) < os.path.getmtime(obj) for d in deps):
        return obj
    cmd = [NVCC] + COMMON + EXTRA.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: compile_one(s, verbose), sources()))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
