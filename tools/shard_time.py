"""Kernel time of one rank's direction slice (world = 1, 2, 4, 8) of the C3 workload on ONE GPU, and a
bitwise check of every slice against the full-grid launch.  usage: python tools/shard_time.py [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
import torch  # noqa: E402
from interface import config  # noqa: E402
from lib import _native, directions  # noqa: E402
from lib.sharded import shard_bounds  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
config.reload(MAX_RES_X=180, MAX_RES_Y=180)
_native.configure_from(config)
L = _native.lib()
M, N, D = config.N_MICROPHONES, config.N_SAMPLES, config.MAX_RES_X * config.MAX_RES_Y
mics, n = directions.active_microphones()
mics = _native.i32(mics)
whole = _native.i32(directions.calculate_delays().astype(int)).ravel()
L.load_coefficients_pad(_native.ptr(whole), whole.size)
_native.check()
gen = torch.Generator(device="cuda").manual_seed(5)
sig = 0.1 * torch.randn((F, M, N), generator=gen, device="cuda")
d_mics = torch.from_numpy(mics).cuda()
st = torch.cuda.current_stream().cuda_stream
full = torch.zeros((D, F), device="cuda")
_native.check(L.bf_mimo_dev_ex(0, sig.data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n, 0, D, 1, F, 0, st))
torch.cuda.synchronize()
for world in (1, 2, 4, 8):
    per, b, c = shard_bounds(D, world, world - 1)
    worst = 0.0
    ok = True
    for rank in range(world):
        per, b, c = shard_bounds(D, world, rank)
        out = torch.zeros((per * world, F), device="cuda")
        call = lambda: _native.check(L.bf_mimo_dev_ex(0, sig.data_ptr(), out.data_ptr(), F, d_mics.data_ptr(), n, b, c, 1, F, 0, st))
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record()
        torch.cuda.synchronize()
        worst = max(worst, e0.elapsed_time(e1) / 5)
        ok = ok and torch.equal(out[b:b + c], full[b:b + c])
        if world == 8 and rank > 1:
            break
    print("world %d  slice kernel %.3f ms  (ideal %.3f)  slices bit-equal to the full launch: %s" % (world, worst, 0, ok))
