"""CPU: the oracle restatement against the REAL reference build (oracle/_ref), bitwise, on
fresh random inputs.  Skipped where oracle/_ref is absent (it is built in the container that
has /root/reference and travels to the GPU box as built files)."""
import numpy as np
import pytest

from oracle import cpu, directions_np as dn, ref
from util import CASES, bits_equal, oracle_cfg

pytestmark = pytest.mark.skipif(not ref.available("default"), reason="oracle/_ref not built")


@pytest.mark.parametrize("case", [c for c in CASES if c != "c3"])
def test_random_inputs_bitwise(case):
    R, dm, cfg = ref.RefC(case), ref.directions(case), oracle_cfg(case)
    mics, n = dm.active_microphones()
    delays = dm.calculate_delays()
    assert bits_equal(delays, dn.calculate_delays(cfg))
    rng = np.random.default_rng(hash(case) % 1000)
    sig = rng.standard_normal((R.M, R.N)).astype(np.float32)
    # geometry table and a random table (delays anywhere in [0, N])
    for whole, d32 in ((delays.astype(int).astype(np.int32), np.float32(delays)),
                       (rng.integers(0, R.N + 1, delays.shape).astype(np.int32),
                        (rng.random(delays.shape) * (R.N - 1)).astype(np.float32))):
        assert bits_equal(R.mimo_pad(sig, mics, whole), cpu.mimo_pad(sig, mics, whole, R.D))
        assert bits_equal(R.mimo_lerp(sig, mics, d32), cpu.mimo_lerp(sig, mics, d32, R.D))
        w1, f1 = R.lerp_tables(d32)
        w2, f2 = cpu.split_lerp(d32)
        assert bits_equal(w1, w2) and bits_equal(f1, f2)
        off = (R.D // 2) * n
        assert bits_equal(R.miso_pad(sig, mics, whole, off), cpu.miso_pad(sig, mics, whole, off))
        assert bits_equal(R.miso_lerp(sig, mics, d32, off), cpu.miso_lerp(sig, mics, d32, off))
    by_mic = rng.integers(0, 40, R.M).astype(np.int32)
    assert bits_equal(R.miso_pad2(sig, mics, by_mic), cpu.miso_pad2(sig, mics, by_mic))
    if R.D <= 400:
        taps = rng.standard_normal(R.D * n * R.T).astype(np.float32)
        for lanes in (0, 1):
            assert bits_equal(R.mimo_fir(sig, mics, taps, lanes), cpu.mimo_fir(sig, mics, taps, R.D, R.T, lanes))
        d32 = np.float32(delays)
        w1, t1 = R.hybrid_tables(d32)
        w2, t2 = cpu.split_hybrid(d32, R.T)
        assert bits_equal(w1, w2) and bits_equal(t1, t2)
        assert bits_equal(R.mimo_hybrid(sig, mics, d32), cpu.mimo_hybrid(sig, mics, d32, R.D, R.T))
        assert bits_equal(dm.compute_convolve_h(), dn.compute_convolve_h(cfg))
