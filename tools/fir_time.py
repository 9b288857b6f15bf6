#!/usr/bin/env python3
"""Time the FIR / hybrid power-map kernels (general one-thread-per-sample kernels) at the stock
config (57x32 grid, 256 mics, 8 taps) through the host-pointer C ABI and on the device."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
import torch  # noqa: E402
from interface import config  # noqa: E402
from lib import _native as nat, directions  # noqa: E402

config.reload()
L = nat.lib()
nat.configure_from(config)
D, n, N, T = 57 * 32, 256, 256, 8
rng = np.random.default_rng(0)
sig = rng.standard_normal((256, 256)).astype(np.float32)
mics = np.arange(256, dtype=np.int32)
taps = nat.f32(directions.compute_convolve_h())
L.load_coefficients_convolve(nat.ptr(taps), taps.size)
_, d32 = directions.whole_and_f32()
L.load_coefficients_convolve_hybrid(nat.ptr(d32), d32.size)
nat.check()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
d_sig = torch.from_numpy(np.repeat(sig[None], F, axis=0)).cuda().contiguous()
d_mics = torch.from_numpy(mics).cuda()
d_img = torch.zeros((F, D), device="cuda")
for simple in (0, 1):
  L.bf_set_kernel_options(simple, 1)
  print("simple kernels" if simple else "tiled kernels", "(frames per launch: %d)" % F)
  for name, algo in (("fir_seq", nat.ALGO_FIR_SEQ), ("fir_lanes", nat.ALGO_FIR_LANES), ("hybrid", nat.ALGO_HYBRID)):
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_mimo_dev(algo, d_sig.data_ptr(), d_img.data_ptr(), F, d_mics.data_ptr(), n, 0, D, None))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts[2:])) / F
    print("  %-10s %.4f ms/map  %.1f maps/s  %.1f GFMA/s (D*n*N*T)" % (name, ms, 1e3 / ms, D * n * N * T / ms / 1e6))
