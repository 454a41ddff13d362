"""Build libptts_b200.so in-tree with nvcc for sm_100a (no torch / pybind dependency: pure C ABI).

Usage: python parler_tts_b200/csrc/build.py [--force]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "gemm.cu", "gemm_tc.cu", "attention.cu", "embed.cu", "sample.cu", "dac.cu", "dac_tc.cu", "step.cu"]
LIB = os.path.join(HERE, "libptts_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
if os.environ.get("PTTS_PTXAS_V"):
    FLAGS += ["-Xptxas", "-v"]


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(HERE)) + ["../../include/ptts_b200.h"]:
        p = os.path.join(HERE, f)
        if f.endswith((".cu", ".cuh", ".h", ".py")) and os.path.isfile(p):
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    stamp_file = os.path.join(HERE, "build", "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)

    def cc(src):
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for src, obj, r in ex.map(cc, SOURCES):
            if verbose and (r.stderr.strip() or r.stdout.strip()):
                print(f"--- {src}\n{r.stdout}{r.stderr}", file=sys.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            objs.append(obj)
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
