#!/usr/bin/env python3
"""MISO stream throughput per delay algorithm (64 mics, 2^15 resident blocks): pad / lerp go through
miso_stream_kernel, FIR / hybrid through the general per-sample kernel."""
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
nat.configure_from(config)
L = nat.lib()
directions.load_pad_from_geometry()
directions.load_lerp_from_geometry()
taps = nat.f32(directions.compute_convolve_h())
L.load_coefficients_convolve(nat.ptr(taps), taps.size)
_, d32 = directions.whole_and_f32()
L.load_coefficients_convolve_hybrid(nat.ptr(d32), d32.size)
nat.check()
mics, n = directions.active_microphones()
d_mics = torch.from_numpy(nat.i32(mics)).cuda()
blocks, M, N = 1 << 15, 64, 256
sig = torch.randn((blocks, M, N), device="cuda")
out = torch.zeros((blocks, N), device="cuda")
off = (14 * 20 + 6) * n
for name, algo in (("pad", nat.ALGO_PAD), ("lerp", nat.ALGO_LERP), ("fir_seq", nat.ALGO_FIR_SEQ),
                   ("fir_lanes", nat.ALGO_FIR_LANES), ("hybrid", nat.ALGO_HYBRID)):
    ts = []
    off_a = off * 8 if algo in (nat.ALGO_FIR_SEQ, nat.ALGO_FIR_LANES) else off     # FIR offsets are tap-table floats
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_miso_dev(algo, sig.data_ptr(), out.data_ptr(), blocks, d_mics.data_ptr(), n, off_a, 1, None))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts[2:]))
    print("%-10s %.3f ms  %.0f GB/s  %.0fx real time" % (name, ms, blocks * (n * N * 4 + N * 4) / ms / 1e6, blocks * N / (ms * 1e-3) / 48828))
