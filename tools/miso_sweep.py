#!/usr/bin/env python3
"""GPU experiment: MISO stream kernel (BASELINE config C2) under different staging choices.
Prints achieved algorithmic GB/s per variant.  Knobs are read by das_miso.cu at launch."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
import torch  # noqa: E402
from interface import config  # noqa: E402
from lib import _native as nat, directions  # noqa: E402

config.reload(N_MICROPHONES=64, N_SAMPLES=256, MAX_RES_X=20, MAX_RES_Y=20, N_TAPS=8, SKIP_N_MICS=1,
              GEOMETRY_N_MICS=64, GEOMETRY_N_ARRAYS=1)
L = nat.lib()
nat.configure_from(config)
directions.load_pad_from_geometry()
directions.load_lerp_from_geometry()
mics, n = directions.active_microphones()
d_mics = torch.from_numpy(nat.i32(mics)).cuda()
blocks, M, N = 1 << 15, 64, 256
sig = torch.randn((blocks, M, N), device="cuda")
out = torch.zeros((blocks, N), device="cuda")
off = (14 * 20 + 6) * n
blk_bytes = n * N * 4 + 2 * n * 4 + N * 4
stream = torch.cuda.current_stream().cuda_stream


def run(algo):
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_miso_dev(algo, sig.data_ptr(), out.data_ptr(), blocks, d_mics.data_ptr(), n, off, 1, stream))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts[2:]))
    return blocks * blk_bytes / (ms * 1e-3) / 1e9


L.bf_set_kernel_options(1, 1)
print("simple kernel (no staging): pad %.0f GB/s  lerp %.0f GB/s" % (run(nat.ALGO_PAD), run(nat.ALGO_LERP)))
L.bf_set_kernel_options(0, 1)
os.environ["BF_MISO_MODE"] = "direct"
print("direct kernel (smem table, 8 loads in flight): pad %.0f GB/s  lerp %.0f GB/s" % (run(nat.ALGO_PAD), run(nat.ALGO_LERP)))
os.environ["BF_MISO_MODE"] = "tma"
for ctas, mt, st, cr in itertools.product((1, 2, 3, 4), (16, 32), (2, 3, 4, 7), (1, 32)):
    if cr > mt:
        continue
    os.environ["BF_MISO_CTAS"] = str(ctas)
    os.environ["BF_MISO_MT"], os.environ["BF_MISO_STAGES"], os.environ["BF_MISO_COPY_ROWS"] = str(mt), str(st), str(cr)
    print("ctas %d Mt %2d stages %d copy_rows %2d: pad %.0f GB/s  lerp %.0f GB/s" % (ctas, mt, st, cr, run(nat.ALGO_PAD), run(nat.ALGO_LERP)), flush=True)
