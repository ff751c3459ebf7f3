#!/usr/bin/env python
"""Install the UNMODIFIED reference Python (the part of Karthik-Ragunath/RAD-NeRF that the hot path touches) under the
git-ignored `baseline/_ref/reference/`, so that it travels to the GPU box with `gpurun` / the driver's push
(`/root/reference` does not exist there).

    python baseline/install_ref.py            # copy from /root/reference (this container)

`pip install --target baseline/_ref /root/reference` is not possible: the reference has neither setup.py nor pyproject.toml
at its root ("Directory '/root/reference' is not installable", recorded in DESIGN.md); its four CUDA extensions have
their own setup.py each and are built by oracle/build_ref.py into oracle/_ref/*.so.  What is copied here, byte for byte:

    nerf/*.py  encoding.py  activation.py  main.py  test.py
    {gridencoder,raymarching,freqencoder,shencoder}/*.py        (the wrappers; their compiled halves are oracle/_ref)

Nothing under baseline/_ref is tracked by git, imported by the product (rad-nerf_b200/), or edited.  Users:
baseline/stock.py (tests and `bench.py --impl reference`).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference")
SRC = os.environ.get("RADNERF_REFERENCE", "/root/reference")
PACKAGES = ("nerf", "gridencoder", "raymarching", "freqencoder", "shencoder")
FILES = ("encoding.py", "activation.py", "main.py", "test.py")


def install(src=SRC, dst=DST):
    if not os.path.isdir(src):
        raise SystemExit("reference tree not found at %s" % src)
    manifest = {}
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    todo = [(f, f) for f in FILES]
    for pkg in PACKAGES:
        for name in sorted(os.listdir(os.path.join(src, pkg))):
            if name.endswith(".py"):
                todo.append((os.path.join(pkg, name), os.path.join(pkg, name)))
    for rel_src, rel_dst in todo:
        a, b = os.path.join(src, rel_src), os.path.join(dst, rel_dst)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
        manifest[rel_dst] = hashlib.sha256(open(b, "rb").read()).hexdigest()
    json.dump({"source": src, "files": manifest}, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return manifest


def installed(dst=DST):
    return os.path.exists(os.path.join(dst, "nerf", "network.py"))


def ensure():
    """install when possible (this container); on the GPU box the copy arrived with the snapshot"""
    if not installed() and os.path.isdir(SRC):
        install()
    return installed()


if __name__ == "__main__":
    m = install()
    print("installed %d files under %s" % (len(m), DST))
    sys.exit(0)
