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


OBJ_DIR = os.path.join(HERE, "build")


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "bf_b200.h"))
    return deps


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def build(force=False, verbose=False):
    """Compile every .cu to its own object (in parallel, only the stale ones), then link."""
    if not force and not stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in _headers())
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)):
            return obj
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        print("+", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    # -z defs: an undefined symbol is a link error here, not a dlopen failure on the GPU box
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xlinker", "-z,defs", *objs, "-o", OUT]
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
