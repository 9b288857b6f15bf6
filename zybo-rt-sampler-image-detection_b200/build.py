#!/usr/bin/env python3
"""Build libbf_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python zybo-rt-sampler-image-detection_b200/build.py [--force] [--verbose]

The .so lands in zybo-rt-sampler-image-detection_b200/lib/ (git-ignored, shipped
to the GPU box with the working tree).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libbf_b200.so")
SOURCES = ["bf_api.cu", "bf_tables.cu", "das_mimo.cu", "das_simple.cu", "das_miso.cu", "das_fir.cu"]
OPTIONAL = ["fd_path.cu", "fd_mvdr.cu", "fd_tc.cu", "ingest.cu", "heatmap.cu", "kf_host.cu", "peer_gather.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
    "--fmad=false",          # every fused multiply-add is written explicitly (bit-exact parity)
]


def sources():
    src = [os.path.join(CSRC, s) for s in SOURCES]
    src += [os.path.join(CSRC, s) for s in OPTIONAL if os.path.exists(os.path.join(CSRC, s))]
    return src


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "include", "bf_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *sources(), "-o", OUT]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
