"""CPU: the N > 1 path (direction sharding + one in-place all-gather) with world_size = 2 over
gloo.  The per-rank slice is computed by the oracle here (tests may use it); on GPUs the same
ShardedMaps object is filled by one bf_mimo_dev_ex launch per rank (bench.py, test_gpu_*)."""
import os
import socket
import sys

import numpy as np
import pytest

from util import ROOT, gold


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from lib.sharded import ShardedMaps
    from oracle import cpu
    from util import gold
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = gold("ragged")                                  # D = 77: not divisible by 2
    delays = g["delays"].reshape(77, -1)
    whole = delays.astype(int).astype(np.int32)
    n = whole.shape[1]
    F = 3
    rng = np.random.default_rng(3)
    frames = rng.standard_normal((F, 256, 256)).astype(np.float32)
    sm = ShardedMaps(77, F, rank, world, "cpu", dist)
    rows = sm.my_rows()
    for f in range(F):
        full = cpu.mimo_pad(frames[f], g["mic_ids"], whole, 77)      # stand-in for the kernel
        rows[:sm.d_count, f] = torch.from_numpy(full[sm.d_begin:sm.d_begin + sm.d_count])
    maps = sm.gather()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), maps.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_grid_exactly():
    sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
    from lib.sharded import shard_bounds
    for D in (1, 7, 77, 400, 1824, 32400):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(D, int)
            for r in range(world):
                per, b, c = shard_bounds(D, world, r)
                assert per * world >= D and 0 <= c <= per
                seen[b:b + c] += 1
            assert np.all(seen == 1)
    assert shard_bounds(32400, 8, 3) == (4050, 12150, 4050)


def test_two_rank_gather_assembles_the_map(tmp_path):
    import torch.multiprocessing as tmp
    from oracle import cpu
    port = _free_port()
    tmp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = gold("ragged")
    whole = g["delays"].reshape(77, -1).astype(int).astype(np.int32)
    rng = np.random.default_rng(3)
    frames = rng.standard_normal((3, 256, 256)).astype(np.float32)
    want = np.stack([cpu.mimo_pad(frames[f], g["mic_ids"], whole, 77) for f in range(3)], axis=1)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == (77, 3) and np.array_equal(got, want)      # every rank holds the full map


def test_peer_gather_layout_model():
    """Layout of the fused kernel + all-gather path (lib.sharded.PeerGather): every rank writes its slice of a
    [world][F][per] buffer on every rank; assemble_peer_layout turns that into [F][D] maps.  NumPy model of
    what bf_mimo_dev_gather stores (frame stride per, direction stride 1, origin d_begin, slice offset
    rank*F*per), for ragged direction counts."""
    import numpy as np
    from lib.sharded import assemble_peer_layout, shard_bounds
    rng = np.random.default_rng(0)
    for D, world, F in ((400, 2, 3), (32400, 8, 2), (91, 4, 5), (7, 8, 1)):
        maps = rng.standard_normal((F, D)).astype(np.float32)
        per = shard_bounds(D, world, 0)[0]
        buf = np.zeros((world, F, per), np.float32)
        flat = buf.reshape(-1)
        for rank in range(world):
            _, d_begin, d_count = shard_bounds(D, world, rank)
            for f in range(F):
                for d in range(d_begin, d_begin + d_count):
                    flat[rank * F * per + f * per + (d - d_begin)] = maps[f, d]      # what the epilogue stores
        assert np.array_equal(assemble_peer_layout(buf, D), maps)


def test_weighted_bounds_tile_the_grid():
    """Slices sized by measured per-GPU speed (lib.sharded.weighted_bounds): contiguous, in rank order, whole
    8-direction groups except at the end of the grid, proportional to the weights, equal weights -> equal slices."""
    from lib.sharded import weighted_bounds
    for D, w in ((32400, [1 / 1.244, 1 / 1.276, 1 / 1.265, 1 / 1.291, 1 / 1.265, 1 / 1.279, 1 / 1.251, 1 / 1.280]),
                 (400, [1, 1, 1]), (91, [3, 1]), (8, [1, 1, 1, 1])):
        b = weighted_bounds(D, w)
        assert len(b) == len(w) and b[0][0] == 0 and sum(c for _, c in b) == D
        for r in range(len(b) - 1):
            assert b[r][0] + b[r][1] == b[r + 1][0] and b[r][1] >= 0 and b[r + 1][0] % 8 == 0
        share = np.array([c for _, c in b]) / D
        assert np.all(np.abs(share - np.array(w) / np.sum(w)) <= 8 / D + 1e-9)
    eq = weighted_bounds(32400, [1.0] * 8)
    assert max(c for _, c in eq) - min(c for _, c in eq) <= 8


def test_gather_layout_rows_are_32_byte_aligned():
    """Row stride of the fused gather buffers: the largest slice rounded up to 8 directions, whatever the slicing
    (a warp's 8 results are one 32-byte peer store), and the slices must tile the grid in rank order."""
    import pytest
    from lib.sharded import gather_layout, shard_bounds, weighted_bounds
    for D, world in ((32400, 8), (32400, 2), (400, 2), (169, 3), (1824, 8), (7, 4)):
        bounds, per = gather_layout(D, world)
        assert bounds == [shard_bounds(D, world, r)[1:] for r in range(world)]
        assert per % 8 == 0 and per >= max(c for _, c in bounds) and per - max(c for _, c in bounds) < 8
        assert sum(c for _, c in bounds) == D
    assert gather_layout(32400, 8)[1] == 4056                       # 4050 directions per rank -> rows of 4056 floats
    wb = weighted_bounds(32400, [1.0, 1.01, 0.99, 1.0, 1.0, 1.0, 1.02, 0.98])
    bounds, per = gather_layout(32400, 8, wb)
    assert bounds == wb and per % 8 == 0 and per >= max(c for _, c in wb)
    ragged, per = gather_layout(400, 2, [(0, 203), (203, 197)])
    assert per == 208
    for bad in ([(0, 200), (201, 199)], [(0, 200)], [(1, 199), (200, 200)], [(0, 300), (300, 101)]):
        with pytest.raises(ValueError):
            gather_layout(400, 2, bad)
