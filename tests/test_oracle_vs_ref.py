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


@pytest.mark.skipif(not ref.RefReceiver.available("default"), reason="oracle/_ref/*/librecv.so not built")
@pytest.mark.parametrize("case,n_arrays", [("default", 4), ("default", 2), ("c1", 1)])
def test_ingest_restatement_vs_the_real_receiver(case, n_arrays):
    """orc_ingest (the oracle the CUDA ingest kernel is tested against) == the reference's own
    receive_and_write_to_buffer / receive_to_buffer (PC/src/receiver.c:94-151, 161-215), fed through a
    socketpair, bit for bit -- including the off-by-one of the odd rows (quirk = 1).  The one read the
    reference makes PAST the payload (last odd row of the last array when the arrays fill the datagram,
    receiver.c:140 with x = 0) lands in an int the test owns: whatever sits there shows up in that one
    channel; the restatement (and the device) read 0 there."""
    import ctypes
    R = ref.RefReceiver(case)
    rng = np.random.default_rng(40 + n_arrays)
    stream = rng.integers(-2 ** 23, 2 ** 23, (R.N, R.M), dtype=np.int32)
    want = np.zeros((R.M, R.N), np.float32)
    cpu.lib().orc_ingest(stream.ctypes.data_as(ctypes.c_void_p), want.ctypes.data_as(ctypes.c_void_p), R.N, R.M,
                         n_arrays, 8, 8, ctypes.c_double(2.0 ** 24), 1)
    n_ch = n_arrays * 64
    for ring in (True, False):
        got = R.receive(stream, n_arrays, counter0=77, past_end=0, ring=ring)
        assert bits_equal(got[:n_ch], want[:n_ch]), (case, n_arrays, ring)
    past = R.receive(stream, n_arrays, past_end=4242)
    differ = np.flatnonzero((past[:n_ch].view(np.uint32) != want[:n_ch].view(np.uint32)).any(axis=1))
    if n_ch == R.M:          # the arrays fill the datagram: channel n_ch - 8 reads stream[N_MICROPHONES]
        assert list(differ) == [n_ch - 8] and np.all(past[n_ch - 8] == np.float32(4242 / 2.0 ** 24))
    else:                    # otherwise the "next row" is still inside the payload
        assert len(differ) == 0
    # the fixed mapping (quirk = 0) differs from the reference exactly on the odd rows
    fixed = np.zeros((R.M, R.N), np.float32)
    cpu.lib().orc_ingest(stream.ctypes.data_as(ctypes.c_void_p), fixed.ctypes.data_as(ctypes.c_void_p), R.N, R.M,
                         n_arrays, 8, 8, ctypes.c_double(2.0 ** 24), 0)
    rows = (np.arange(n_ch) // 8) % 8
    assert bits_equal(fixed[:n_ch][rows % 2 == 0], want[:n_ch][rows % 2 == 0])
    assert not bits_equal(fixed[:n_ch][rows % 2 == 1], want[:n_ch][rows % 2 == 1])


@pytest.mark.skipif(not ref.v4_available("c1"), reason="no AVX-512 on this host or libref_v4.so not built")
def test_avx512_build_of_the_reference_is_bitwise_the_portable_one():
    """bench.py's reference arm may time libref_v4.so (-march=x86-64-v4, what the reference's -march=native gives
    on an AVX-512 host) beside libref.so (-mavx2 -mfma): same source, and the outputs must be the same bits."""
    a, b = ref.RefC("c1"), ref.RefC("c1", "libref_v4.so")
    rng = np.random.default_rng(3)
    sig = rng.standard_normal((a.M, a.N)).astype(np.float32)
    mics = np.arange(a.M, dtype=np.int32)
    whole = rng.integers(0, 48, (a.D, a.M)).astype(np.int32)
    d32 = (rng.random((a.D, a.M)) * 47.5).astype(np.float32)
    assert bits_equal(a.mimo_pad(sig, mics, whole), b.mimo_pad(sig, mics, whole))
    assert bits_equal(a.mimo_lerp(sig, mics, d32), b.mimo_lerp(sig, mics, d32))
