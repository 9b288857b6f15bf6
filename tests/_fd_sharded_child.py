"""Child of test_gpu_peer_gather.py::test_fd_sharded_two_ranks: run under torchrun with 2+ ranks, one GPU each.
Direction-sharded frequency-domain maps (FD path): every rank steers its slice of the grid
(bf_fd_mvdr_dev_slice / bf_fd_das_dev_slice), the slices travel through lib.sharded.PeerGather (NVLink peer
stores), and every rank's assembled map is compared with a one-GPU run over all directions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    from lib import _native as nat
    from lib.sharded import PeerGather, fd_das_sharded, fd_mvdr_sharded, fd_mvdr_sharded_bins
    import realtime_scripts.calc_r_prime as rp
    import realtime_scripts.config as cfg
    L = nat.lib()
    L.bf_set_device(int(os.environ["LOCAL_RANK"]))
    M, N, K, RES_X, RES_Y = 256, 256, 8, 33, 31                    # 1023 directions: not a multiple of the MMA tile
    D = RES_X * RES_Y
    pos, _ = rp.calc_r_prime(cfg.ELEMENT_DISTANCE)
    x_max = np.tan(np.deg2rad(cfg.VIEW_ANGLE / 2))
    xs = np.linspace(-x_max, x_max, RES_X)
    ys = np.linspace(-x_max / cfg.ASPECT_RATIO, x_max / cfg.ASPECT_RATIO, RES_Y)
    mx, my = np.ascontiguousarray(pos[0]), np.ascontiguousarray(pos[1])
    act = np.arange(M, dtype=np.int32)
    p = nat.ptr
    nat.check(L.bf_fd_setup(M, N, 48828.0, 343.0, 4, 60, p(xs), RES_X, p(ys), RES_Y, 1.0, p(mx), p(my), p(act), M))
    gen = torch.Generator(device="cuda").manual_seed(11)            # same seed on every rank: same data
    t = torch.arange(N, device="cuda")[None, None, :]
    m = torch.arange(M, device="cuda")[None, :, None]
    snaps = 0.05 * torch.randn((K, M, N), generator=gen, device="cuda")
    snaps += 0.3 * torch.sin(2 * np.pi * 3000.0 * (t + 0.013 * m * 48.828) / 48828.0)
    snaps = snaps.float().contiguous()
    ok = True
    # ---- MVDR ----
    full = torch.zeros(D, device="cuda")
    nat.check(L.bf_fd_mvdr_dev(snaps.data_ptr(), full.data_ptr(), K, 1e-2, None))
    torch.cuda.synchronize()
    pg = PeerGather(D, 1, rank, world, dist, depth=2)
    for i in range(3):
        fd_mvdr_sharded(pg, i, snaps, K, 1e-2)
        got = pg.maps(i).clone().reshape(-1)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(full, got))
        if not torch.equal(full, got):
            print("rank %d mvdr step %d max rel diff %.3e" % (rank, i, float(((full - got).abs() / full).max())))
    # the float64 stages sharded by bins as well (operand images all-gathered): same map
    n_bins = 56
    for i in range(3, 5):
        fd_mvdr_sharded_bins(pg, i, snaps, K, 1e-2, n_bins, dist)
        got = pg.maps(i).clone().reshape(-1)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(full, got))
        if not torch.equal(full, got):
            print("rank %d mvdr (bins sharded) step %d max rel diff %.3e" % (rank, i, float(((full - got).abs() / full).max())))
    pg.check()
    dist.barrier()
    pg.close()
    # ---- DAS, 3 frames ----
    F = 3
    frames = snaps[:F].contiguous()
    ref = torch.zeros((F, D), device="cuda")
    nat.check(L.bf_fd_das_dev(frames.data_ptr(), ref.data_ptr(), F, 0.2, 1, None))
    torch.cuda.synchronize()
    pg = PeerGather(D, F, rank, world, dist, depth=2)
    for i in range(2):
        got = fd_das_sharded(pg, i, frames, 0.2, True)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(ref, got))
    pg.check()
    dist.barrier()
    pg.close()
    tt = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("FD_SHARDED_OK" if int(tt) else "FD_SHARDED_MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
