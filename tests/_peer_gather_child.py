"""Child of test_gpu_peer_gather.py: run under torchrun with 2+ ranks, one GPU each.  Computes C1 power maps
with the fused kernel + NVLink peer-store all-gather (lib.sharded.PeerGather) for several steps and compares
every rank's assembled maps bit-for-bit with a one-GPU launch over all directions."""
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
    from util import product_config
    from lib import _native as nat, directions
    from lib.sharded import PeerGather
    config = product_config("c1")
    nat.configure_from(config)
    L = nat.lib()
    L.bf_set_device(int(os.environ["LOCAL_RANK"]))
    M, N, D = config.N_MICROPHONES, config.N_SAMPLES, config.MAX_RES_X * config.MAX_RES_Y
    mics, n = directions.active_microphones()
    mics = nat.i32(mics)
    whole, d32 = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    F, steps = 5, 7
    gen = torch.Generator(device="cuda").manual_seed(3)            # same seed on every rank: same signals
    sig = 0.1 * torch.randn((steps, F, M, N), generator=gen, device="cuda")
    d_mics = torch.from_numpy(mics).cuda()
    ok = True
    # third pass: pad with overlapping steps (programmatic stream serialisation, bf_gather_overlap); fourth: ragged
    # slices whose sizes are not multiples of the 8-direction group (the buffers' row stride is padded)
    ragged = [(0, D // world + 3)] + [(D // world + 3 + (r - 1) * ((D - D // world - 3) // (world - 1)),
                                       (D - D // world - 3) // (world - 1)) for r in range(1, world)]
    ragged[-1] = (ragged[-1][0], D - ragged[-1][0])
    for algo, overlap, bounds in ((nat.ALGO_PAD, False, None), (nat.ALGO_LERP, False, None), (nat.ALGO_PAD, True, None),
                                  (nat.ALGO_PAD, True, ragged)):
        pg = PeerGather(D, F, rank, world, dist, depth=3, consume_lag=1, overlap=overlap, bounds=bounds)
        assert pg.per % 8 == 0
        got = []
        for i in range(steps):
            pg.step(i, algo, sig[i], d_mics, n)
            if i >= 1:
                got.append(pg.maps(i - 1).clone())                 # consume step i-1 while step i is in flight
        got.append(pg.maps(steps - 1).clone())
        torch.cuda.synchronize()
        pg.check()
        for i in range(steps):
            full = torch.zeros((F, D), device="cuda")
            nat.check(L.bf_mimo_dev(algo, sig[i].data_ptr(), full.data_ptr(), F, d_mics.data_ptr(), n, 0, D, None))
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(full, got[i]))
        dist.barrier()
        pg.close()
    # input all-gather on the copy engines (PeerInput): every rank pushes its frames of a step, all ranks end up
    # with the whole batch; slots are reused only after every rank's kernel of the step that used them finished
    from lib.sharded import PeerInput
    Fp = 3
    whole = 0.1 * torch.randn((steps, world * Fp, M, N), generator=gen, device="cuda")      # same on every rank
    h_parts = whole[:, rank * Fp:(rank + 1) * Fp].cpu().pin_memory()
    pin = PeerInput((Fp, M, N), rank, world, dist, slots=2)
    pg = PeerGather(D, world * Fp, rank, world, dist, depth=3, consume_lag=1)
    s_in = torch.cuda.Stream()
    got = []
    for i in range(steps):
        sl = i & 1
        if i >= 2:
            pg.ready(i - 2, s_in.cuda_stream)
        pin.push(sl, h_parts[i], s_in.cuda_stream)
        batch = pin.wait(sl)
        ok = ok and bool(torch.equal(batch, whole[i]))
        pg.step(i, nat.ALGO_PAD, batch, d_mics, n)
        if i >= 1:
            got.append(pg.maps(i - 1).clone())
    got.append(pg.maps(steps - 1).clone())
    torch.cuda.synchronize()
    pin.check()
    pg.check()
    for i in range(steps):
        full = torch.zeros((world * Fp, D), device="cuda")
        nat.check(L.bf_mimo_dev(nat.ALGO_PAD, whole[i].data_ptr(), full.data_ptr(), world * Fp, d_mics.data_ptr(), n, 0, D, None))
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(full, got[i]))
    dist.barrier()
    pin.close()
    pg.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER_GATHER_OK" if int(t) else "PEER_GATHER_MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
