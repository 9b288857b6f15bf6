#!/usr/bin/env python3
"""BASELINE config C4 on one GPU: frequency-domain MVDR, 256 mics, 1024-point FFT, bins 1..512,
K = 64 snapshots, 256 x 128 = 32 768 directions.  Prints per-stage device times and the useful /
issued tensor throughput of the steering contraction (8*D*M^2*F useful flops).

    python tools/mvdr_c4.py [--bins 512] [--dirs 32768] [--snapshots 64] [--reps 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=512)
    ap.add_argument("--dirs", type=int, default=32768)
    ap.add_argument("--snapshots", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tc", type=int, default=4)
    args = ap.parse_args()
    os.environ["BF_MVDR_TC"] = str(args.tc)
    import torch
    from lib import _native as nat
    import realtime_scripts.calc_r_prime as rp
    import realtime_scripts.config as cfg
    L = nat.lib()
    M, N, K, F = 256, 1024, args.snapshots, args.bins
    res_x = 256
    res_y = args.dirs // res_x
    D = res_x * res_y
    pos_all, _ = rp.calc_r_prime(cfg.ELEMENT_DISTANCE)
    x_max = np.tan(np.deg2rad(cfg.VIEW_ANGLE / 2))
    xs = np.linspace(-x_max, x_max, res_x)
    ys = np.linspace(-x_max / cfg.ASPECT_RATIO, x_max / cfg.ASPECT_RATIO, res_y)
    act = np.arange(M, dtype=np.int32)
    p = nat.ptr
    mx, my = np.ascontiguousarray(pos_all[0]), np.ascontiguousarray(pos_all[1])
    nat.check(L.bf_fd_setup(M, N, 48828.0, 343.0, 1, 1 + F, p(xs), res_x, p(ys), res_y, 1.0, p(mx), p(my), p(act), M))
    gen = torch.Generator(device="cuda").manual_seed(1237)
    snaps = 0.05 * torch.randn((K, M, N), generator=gen, device="cuda")
    # three tones, steered by a linear phase ramp across microphones (synthetic, seed 1237)
    t = torch.arange(N, device="cuda")[None, None, :]
    m = torch.arange(M, device="cuda")[None, :, None]
    for f0, amp, slope in ((2000.0, 0.3, 0.011), (5000.0, 0.2, -0.023), (9000.0, 0.1, 0.005)):
        ph = 2 * np.pi * torch.rand((K, 1, 1), generator=gen, device="cuda")
        snaps += amp * torch.sin(2 * np.pi * f0 * (t + slope * m * 48.828) / 48828.0 + ph)
    snaps = snaps.float().contiguous()
    power = torch.zeros(D, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    times = []
    for r in range(args.reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_fd_mvdr_dev(snaps.data_ptr(), power.data_ptr(), K, 1e-2, stream))
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = float(np.mean(times[1:]))
    st = np.zeros(5, np.float32)
    L.bf_fd_mvdr_timings(nat.ptr(st))
    useful = 8.0 * D * M * M * F
    out = {"workload": "C4: FD-MVDR, 256 mics, 1024-pt FFT, %d bins, K=%d, %d directions" % (F, K, D),
           "ms_per_map_all_stages": ms, "maps_per_s": 1e3 / ms,
           "steer_useful_tflop": useful / 1e12, "finite": bool(torch.isfinite(power).all()),
           "power_max": float(power.max()), "tc": args.tc,
           "stage_ms": dict(zip(["fft64", "covariance", "cholesky", "tri_inverse", "steering"], [float(x) for x in st])),
           "steer_useful_tflops": useful / 1e12 / (float(st[4]) * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
