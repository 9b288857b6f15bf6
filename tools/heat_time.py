"""Per-kernel timing of the heat-map stage (csrc/heatmap.cu) at the C3/C5 grid.
usage: python tools/heat_time.py [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
import torch  # noqa: E402
from interface import config  # noqa: E402
from lib import _native, visual  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
config.reload(MAX_RES_X=180, MAX_RES_Y=180)
X = Y = 180
L = _native.lib()
gen = torch.Generator(device="cuda").manual_seed(1)
maps = (torch.rand((frames, X * Y), generator=gen, device="cuda") ** 6) * 1e-3
st = torch.cuda.current_stream().cuda_stream
lut = visual.generate_color_map()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


small = torch.empty((frames, Y, X, 3), dtype=torch.uint8, device="cuda")
index = torch.empty((frames, X, Y), dtype=torch.int16, device="cuda")
info = torch.empty((frames, 48), dtype=torch.uint8, device="cuda")
conf = torch.empty(frames, dtype=torch.float64, device="cuda")
t_heat = timed(lambda: _native.check(L.bf_heatmap_dev(maps.data_ptr(), frames, X * Y, X, Y, 1e-7, 0.5, 5, 1, _native.ptr(lut),
                                                      small.data_ptr(), index.data_ptr(), info.data_ptr(), st)))
print("frames %d  heat_kernel %.3f ms (%.1f us/frame)" % (frames, t_heat, 1e3 * t_heat / frames))
for (W, H) in ((640, 360), (1920, 1080)):
    big = torch.empty((frames, H, W, 3), dtype=torch.uint8, device="cuda")
    t_rs = timed(lambda: _native.check(L.bf_resize_linear_u8_dev(small.data_ptr(), frames, Y, X, 3, big.data_ptr(), H, W, st)))
    t_en = timed(lambda: _native.check(L.bf_entropy_dev(big.data_ptr(), frames, H * W * 3, conf.data_ptr(), st)))
    gb = frames * H * W * 3 / 1e9
    print("  %dx%d  resize %.3f ms (%.0f GB/s written)   entropy %.3f ms (%.0f GB/s read)" % (W, H, t_rs, gb / t_rs * 1e3, t_en, gb / t_en * 1e3))
    del big
