"""Helper for test_producer_loop_in_child_process: runs in a FRESH interpreter (no CUDA in the
parent), forks the producer loop the way the reference's demos do (main.pyx:702-721,
multiprocessing fork start method) and stores what arrives on the queue."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200")):
    sys.path.insert(0, p)


def producer(target_name, q, running, case):
    from util import gold, product_config
    product_config(case)
    from lib import beamformer
    g = gold(case)
    rec = np.concatenate([g["signals"], g["signals"][:, ::-1]], axis=1)
    beamformer.connect(False, verbose=False, source=beamformer.ArraySource(rec))
    getattr(beamformer, target_name)(q, running)


def main(target, out_path):
    ctx = mp.get_context("fork")
    q, running = ctx.JoinableQueue(maxsize=4), ctx.Value("i", 1)
    p = ctx.Process(target=producer, args=(target, q, running, "c1"))
    p.start()
    items = []
    try:
        for _ in range(3):
            items.append(q.get(timeout=180))
    finally:
        running.value = 0
        while True:
            try:
                q.get(timeout=2)
            except Exception:  # noqa: BLE001
                break
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    maps, nrs = [], []
    for it in items:
        if isinstance(it, tuple):
            maps.append(it[0]); nrs.append(it[1])
        else:
            maps.append(it); nrs.append(-1)
    np.savez(out_path, maps=np.stack(maps), nrs=np.array(nrs),
             contiguous=np.array([m.flags["C_CONTIGUOUS"] for m in maps]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
