"""CPU: the launch plan and tile hand-out of the tiled power-map kernel (csrc/das_mimo.cu: plan_launch, TileWalk),
called through the library's host-only entry points bf_mimo_plan / bf_mimo_walk -- the same code the kernel runs.
Whatever the launch shape, every (frame, group) unit is processed exactly once, a tile never crosses a frame or
exceeds the consumer warps, and contiguous ranges give every CTA the same work (+-1 unit)."""
import ctypes

import numpy as np
import pytest

from lib import _native as nat


def _plan(L, lerp, groups, frames, overlap, sms):
    W, grid, ranges = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.bf_mimo_plan(lerp, groups, frames, overlap, sms, ctypes.byref(W), ctypes.byref(grid), ctypes.byref(ranges)) == 0
    return W.value, grid.value, ranges.value


def _walk(L, lerp, groups, frames, overlap, sms, block):
    cap = 4 * (groups * frames // max(1, sms) + groups + frames + 8)
    buf = np.zeros((cap, 3), np.int32)
    k = L.bf_mimo_walk(lerp, groups, frames, overlap, sms, block, nat.ptr(buf), cap)
    assert 0 <= k <= cap
    return buf[:k]


@pytest.mark.parametrize("lerp", [0, 1])
@pytest.mark.parametrize("overlap", [0, 1])
def test_every_unit_exactly_once(lerp, overlap):
    L = nat.lib()
    rng = np.random.default_rng(11 + lerp + 2 * overlap)
    shapes = [(4050, 128, 148), (4050, 1, 148), (507, 128, 148), (507, 16, 148), (228, 1, 148), (50, 37, 148),
              (1, 1, 148), (22, 1, 148), (4050, 16, 132), (3, 500, 148)]
    shapes += [(int(rng.integers(1, 700)), int(rng.integers(1, 40)), int(rng.choice([4, 37, 132, 148]))) for _ in range(12)]
    for groups, frames, sms in shapes:
        W, grid, ranges = _plan(L, lerp, groups, frames, overlap, sms)
        assert 1 <= W <= (15 if lerp else 19) and 1 <= grid <= sms
        if lerp or overlap:
            assert ranges == 0
        seen = np.zeros((frames, groups), np.int32)
        per_block = []
        for b in range(grid):
            tiles = _walk(L, lerp, groups, frames, overlap, sms, b)
            units = 0
            for frame, g0, cnt in tiles:
                assert 0 <= frame < frames and 0 <= g0 and 1 <= cnt <= W and g0 + cnt <= groups, (groups, frames, sms, b)
                seen[frame, g0:g0 + cnt] += 1
                units += cnt
            per_block.append(units)
            if ranges:                      # a CTA's range is contiguous in frame-major unit order
                flat = [f * groups + g for f, g0, c in tiles for g in range(g0, g0 + c)]
                assert flat == list(range(flat[0], flat[0] + len(flat))) if flat else True
        assert np.all(seen == 1), (groups, frames, sms, W, grid, ranges)
        if ranges:
            assert max(per_block) - min(per_block) <= 1
        assert L.bf_mimo_walk(lerp, groups, frames, overlap, sms, grid, None, 0) == -1       # beyond the grid


def test_c3_plans():
    """The shapes the bench runs (148 SMs): what the launcher picks, as documented in DESIGN.md 4.1 / 5."""
    L = nat.lib()
    assert _plan(L, 0, 4050, 128, 0, 148) == (19, 148, 1)          # C3 pad, 128 frames: contiguous ranges
    assert _plan(L, 1, 4050, 128, 0, 148) == (15, 148, 0)          # lerp keeps whole tiles
    assert _plan(L, 0, 4050, 1, 0, 148) == (15, 148, 0)            # a single frame: 15 warps (two rounds either way)
    assert _plan(L, 0, 507, 128, 1, 148) == (19, 148, 0)           # 1/8 slice in an overlapping step: whole tiles
    W, grid, ranges = _plan(L, 0, 228, 1, 0, 148)                  # stock 57 x 32 grid, one frame: spread over the SMs
    assert (W, ranges) == (2, 0) and grid == 114
    assert L.bf_mimo_plan(0, 0, 1, 0, 148, None, None, None) != 0


def test_host_batch_chunk_schedule():
    """bf_mimo_host_batch's chunks: they tile the batch, none exceeds 32 frames, and long batches start and end
    with 4-frame chunks (the two copies the call cannot hide behind a kernel)."""
    L = nat.lib()
    for frames in list(range(1, 200)) + [255, 256, 257, 1000, 4096]:
        buf = np.zeros(256, np.int32)
        k = L.bf_host_batch_schedule(frames, nat.ptr(buf), 256)
        c = buf[:k]
        assert k >= 1 and c.sum() == frames and c.min() >= 1 and c.max() <= 32, (frames, c)
        if frames > 8:
            assert c[0] == 4 and c[-1] == 4
        if frames >= 64:
            assert c[1] == 12 and c[-2] == 12
    buf = np.zeros(8, np.int32)
    assert L.bf_host_batch_schedule(128, nat.ptr(buf), 8) == 7 and list(buf[:7]) == [4, 12, 32, 32, 32, 12, 4]
    assert L.bf_host_batch_schedule(0, None, 0) == -1
