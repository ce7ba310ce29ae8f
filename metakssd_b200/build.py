"""Build libmkssd_b200.so in-tree with nvcc for sm_100a (B200).  No JIT, no torch dependency."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libmkssd_b200.so")
SRC = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
HDR = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SRC + HDR)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SRC:
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC,-Wno-format-truncation,-pthread", "-I" + os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        cmd += os.environ.get("MK_NVCC_FLAGS", "").split()
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    # gcc-13 wrapper in this image needs the system g++ for linking
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart", "-lpthread", "-ldl"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
