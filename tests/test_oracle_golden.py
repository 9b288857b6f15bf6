"""CPU: the oracle restatement (oracle/oracle.c + oracle/directions_np.py) against the golden
vectors generated from the real reference build (oracle/gen_golden.py).  Bit-exact."""
import numpy as np
import pytest

from oracle import cpu, directions_np as dn
from util import CASES, bits_equal, gold, oracle_cfg, sha


@pytest.mark.parametrize("case", list(CASES))
def test_delay_tables_bit_exact(case):
    g, cfg = gold(case), oracle_cfg(case)
    mics, n = dn.active_microphones(cfg)
    assert np.array_equal(mics, g["mic_ids"])
    delays = dn.calculate_delays(cfg)
    assert tuple(g["grid"]) == delays.shape[:2]
    assert sha(delays) == str(g["delays_sha"])                      # float64 table, bitwise
    assert sha(delays.astype(int).astype(np.int32)) == str(g["whole_sha"])
    assert sha(np.float32(delays)) == str(g["d32_sha"])
    assert bits_equal(dn.calc_r_prime(cfg, float(np.float32(0.02))), g["r_prime"])
    whole, weight = cpu.split_lerp(np.float32(delays))
    assert sha(whole) == str(g["lerp_whole_sha"])
    assert sha(weight) == str(g["lerp_weight_sha"])
    if "delays" in g:
        assert bits_equal(delays, g["delays"])


def _signals(case, g):
    if "signals" in g:
        return g["signals"]
    from lib import synthetic
    s = synthetic.plot_py_stimulus(256, 256)
    assert sha(s) == str(g["signals_sha"])
    return s


@pytest.mark.parametrize("case", list(CASES))
def test_power_maps_bit_exact(case):
    g, cfg = gold(case), oracle_cfg(case)
    sig = _signals(case, g)
    delays = dn.calculate_delays(cfg)
    D, n = delays.shape[0] * delays.shape[1], delays.shape[2]
    whole, d32, mics = delays.astype(int), np.float32(delays), g["mic_ids"]
    assert bits_equal(cpu.mimo_pad(sig, mics, whole, D), g["img_pad"])
    assert bits_equal(cpu.mimo_lerp(sig, mics, d32, D), g["img_lerp"])
    for i, d in enumerate(g["miso_dirs"]):
        assert bits_equal(cpu.miso_pad(sig, mics, whole, d * n), g["miso_pad"][i])
        assert bits_equal(cpu.miso_lerp(sig, mics, d32, d * n), g["miso_lerp"][i])
    if "wrapper_pad" in g:   # the reference's own Python wrappers on the plot.py stimulus
        assert bits_equal(g["wrapper_pad"].ravel(), g["img_pad"])
        assert bits_equal(g["wrapper_lerp"].ravel(), g["img_lerp"])
        assert abs(float(g["wrapper_pad"].max()) - 0.5001457) < 1e-6
        assert np.unravel_index(g["wrapper_pad"].argmax(), (57, 32)) == (28, 14)
        assert np.unravel_index(g["wrapper_lerp"].argmax(), (57, 32)) == (28, 16)


@pytest.mark.parametrize("case", ["c1", "ragged", "taps64"])
def test_fir_and_hybrid_bit_exact(case):
    g, cfg = gold(case), oracle_cfg(case)
    sig, mics = g["signals"], g["mic_ids"]
    delays = dn.calculate_delays(cfg)
    D, T = delays.shape[0] * delays.shape[1], cfg["N_TAPS"]
    taps = dn.compute_convolve_h(cfg)
    assert sha(taps) == str(g["taps_sha"])
    assert bits_equal(cpu.mimo_fir(sig, mics, taps, D, T, 0), g["img_fir_seq"])
    assert bits_equal(cpu.mimo_fir(sig, mics, taps, D, T, 1), g["img_fir_lanes"])
    hw, ht = cpu.split_hybrid(np.float32(delays), T)
    assert sha(ht) == str(g["hybrid_taps_sha"])
    assert bits_equal(cpu.mimo_hybrid(sig, mics, np.float32(delays), D, T), g["img_hybrid"])


def test_edge_cases_oracle():
    """Empty / maximal delays, single microphone, ragged sizes (reference semantics)."""
    rng = np.random.default_rng(5)
    N, M = 256, 8
    sig = rng.standard_normal((M, N)).astype(np.float32)
    mics = np.arange(M, dtype=np.int32)
    # delay == N: no contribution at all -> zero power; delay == 0: plain sum
    img = cpu.mimo_pad(sig, mics, np.full((2, M), N, np.int32), 2)
    assert np.all(img == 0)
    img0 = cpu.mimo_pad(sig, mics, np.zeros((1, M), np.int32), 1)
    s = np.zeros(N, np.float32)
    for m in range(M):
        s = s + sig[m]
    ref = np.float32(0)
    for k in range(N):
        x = np.float32(s[k] / np.float32(M))
        ref = np.float32(ref + np.float32(x * x))
    assert img0[0] == np.float32(ref / np.float32(N))
    # lerp with an integer delay: weight 1-0 = 1 -> s[i] + 1*(s[i+1]-s[i]), shifted by w+1
    out = cpu.miso_lerp(sig, mics[:1], np.full(1, 3.0, np.float32), 0)
    assert np.all(out[:4] == 0)
    exp = sig[0, :-4] + np.float32(1.0) * (sig[0, 1:-3] - sig[0, :-4])
    assert bits_equal(out[4:], exp.astype(np.float32))


def test_steer_offset_arithmetic():
    cfg = dn.cfg_with()
    assert dn.steer_offset_degree(cfg, 0, 0, 256) == int(16 * 57 * 256 + 28 * 256)
    assert dn.steer_offset_unit(cfg, 0.5, 0.5, 256) == int(16 * 57 * 256 + 28 * 256)
    assert dn.steer_offset_degree(cfg, -90, -90, 64) == 0
