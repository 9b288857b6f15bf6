#!/usr/bin/env python3
"""Frequency-domain DAS (a19) at the reference's stock size (13x13, 94 bins) and at config C4's size
(512 bins, 32 768 directions): device time per frame of bf_fd_das_dev."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
import torch  # noqa: E402
from lib import _native as nat  # noqa: E402
import realtime_scripts.calc_r_prime as rp  # noqa: E402
import realtime_scripts.config as cfg  # noqa: E402

L = nat.lib()
p = nat.ptr
pos_all, _ = rp.calc_r_prime(cfg.ELEMENT_DISTANCE)
mx, my = np.ascontiguousarray(pos_all[0]), np.ascontiguousarray(pos_all[1])
M = 256
act = np.arange(M, dtype=np.int32)
x_max = np.tan(np.deg2rad(cfg.VIEW_ANGLE / 2))
for (N, lo, hi, rx, ry, frames) in ((256, 0, 94, 13, 13, 64), (1024, 1, 513, 256, 128, 2)):
    xs = np.linspace(-x_max, x_max, rx)
    ys = np.linspace(-x_max / cfg.ASPECT_RATIO, x_max / cfg.ASPECT_RATIO, ry)
    nat.check(L.bf_fd_setup(M, N, 48828.0, 343.0, lo, hi, p(xs), rx, p(ys), ry, 1.0, p(mx), p(my), p(act), M))
    D, F = rx * ry, hi - lo
    gen = torch.Generator(device="cuda").manual_seed(2)
    sig = torch.randn((frames, M, N), generator=gen, device="cuda")           # [frames][mic][sample] (include/bf_b200.h)
    out = torch.zeros((frames, D), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for r in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_fd_das_dev(sig.data_ptr(), out.data_ptr(), frames, 0.2, 1, st))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts[1:])) / frames
    print("N=%d bins=%d dirs=%d: %.3f ms/frame  %.1f frames/s  %.2f TFLOP/s (8*D*M*F)" % (N, F, D, ms, 1e3 / ms, 8.0 * D * M * F / ms / 1e9))
