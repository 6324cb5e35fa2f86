"""Build recipe for libofb.so — explicit nvcc for sm_100a, in-tree output.

``python -m opticalflowcontainer_b200.build`` (or ``__graft_entry__.build()``) compiles every
``csrc/*.cu`` with ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and links them into
``opticalflowcontainer_b200/libofb.so`` (static cudart, no torch types anywhere).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libofb.so")
BUILD = os.path.join(HERE, "csrc", "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math=false"]
CFLAGS += os.environ.get("OFB_NVCC_FLAGS", "").split()   # experiments only (e.g. -DOFB_DBG=1)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    m = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and f.endswith((".cu", ".cuh", ".h")):
            m.update(f.encode()); m.update(open(p, "rb").read())
    m.update(open(os.path.join(HERE, "..", "include", "ofb.h"), "rb").read())
    m.update(" ".join(ARCH + CFLAGS).encode())
    return m.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    flags = [f for f in CFLAGS if not f.startswith("--use_fast_math")]
    objs = []

    def cc(src):
        obj = os.path.join(BUILD, src[:-3] + ".o")
        cmd = [NVCC, *ARCH, *flags, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, _sources()))
    cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", OUT, *objs, "-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
