#!/usr/bin/env python3
"""Run the GATHER instantiation of das_mimo_kernel on ONE GPU (world = 1: own buffer only, step flags on) so that it
can be captured with ncu (ncu must not wrap a multi-rank command).  C3, pad, 128 frames, a 1/8 direction slice --
the launch shape of an 8-GPU step.

    python tools/gather_single.py [--slice 8] [--frames 128] [--reps 5]
    python tools/gather_single.py --run 40 [--overlap 1] [--depth 4]     # 40 back-to-back steps, ms per step
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slice", type=int, default=8)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--run", type=int, default=0, help="time this many back-to-back steps instead of single launches")
    ap.add_argument("--overlap", type=int, default=0, help="bf_gather_overlap for the run")
    ap.add_argument("--depth", type=int, default=4)
    args = ap.parse_args()
    import torch
    from interface import config
    config.reload(N_MICROPHONES=256, N_SAMPLES=256, MAX_RES_X=180, MAX_RES_Y=180, N_TAPS=8, SKIP_N_MICS=1,
                  GEOMETRY_N_MICS=256, GEOMETRY_N_ARRAYS=4)
    from lib import _native as nat, directions
    L = nat.lib()
    nat.configure_from(config)
    directions.load_pad_from_geometry()
    mics, n = directions.active_microphones()
    d_mics = torch.from_numpy(nat.i32(mics)).cuda()
    D, F = 180 * 180, args.frames
    per = (D + args.slice - 1) // args.slice
    sig = 0.1 * torch.randn((F, 256, 256), device="cuda")
    buf = torch.zeros((1, F, per), device="cuda")
    flags = torch.zeros(8, dtype=torch.int64, device="cuda")
    timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    vp = ctypes.c_void_p
    bufs = (vp * 1)(vp(buf.data_ptr()))
    fl = (vp * 1)(vp(flags.data_ptr()))
    st = torch.cuda.current_stream().cuda_stream
    if args.run:
        ring = [torch.zeros((1, F, per), device="cuda") for _ in range(args.depth)]
        pool = [0.1 * torch.randn((F, 256, 256), device="cuda") for _ in range(3)]
        out = {}
        seq = 0
        for name, ov in (("plain", 0), ("overlap", 1)) if args.overlap else (("plain", 0),):
            for rep in range(3):
                L.bf_gather_overlap(ov)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for i in range(args.run):
                    seq += 1
                    bufs_i = (vp * 1)(vp(ring[seq % args.depth].data_ptr()))
                    nat.check(L.bf_mimo_dev_gather_sync(nat.ALGO_PAD, pool[i % 3].data_ptr(), F, d_mics.data_ptr(), n, 0, per,
                                                        0, 1, bufs_i, per, fl, max(0, seq - args.depth + 1), seq,
                                                        timed_out.data_ptr(), st))
                b.record()
                torch.cuda.synchronize()
                L.bf_gather_overlap(0)
                out[name] = a.elapsed_time(b) / args.run
            last = (args.run - 1) % 3
            full = torch.zeros((F, D), device="cuda")
            nat.check(L.bf_mimo_dev(nat.ALGO_PAD, pool[last].data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n, 0, D, None))
            torch.cuda.synchronize()
            out[name + "_last_step_matches"] = bool(torch.equal(full[:, :per], ring[seq % args.depth][0]))
        out.update(frames=F, slice=args.slice, timed_out=int(timed_out), flags=int(flags[0]))
        print(out)
        return
    ts = []
    for i in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(L.bf_mimo_dev_gather_sync(nat.ALGO_PAD, sig.data_ptr(), F, d_mics.data_ptr(), n, 0, per, 0, 1, bufs, per,
                                            fl, i, i + 1, timed_out.data_ptr(), st))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    full = torch.zeros((F, D), device="cuda")
    nat.check(L.bf_mimo_dev(nat.ALGO_PAD, sig.data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n, 0, D, None))
    torch.cuda.synchronize()
    print({"kernel_ms": float(np.mean(ts[1:])), "flags": int(flags[0]), "timed_out": int(timed_out),
           "matches_full_launch": bool(torch.equal(full[:, :per], buf[0]))})


if __name__ == "__main__":
    main()
