"""GPU: the CUDA path (through the C ABI of libbf_b200.so) against the golden vectors of the
real reference and against the oracle on seeded inputs.

Bars: integer tables, sample indexing and -- because the microphone sum and the epilogue are
evaluated in the reference's order -- the fp32 power maps of pad/lerp/FIR/hybrid are required to
be BIT-EXACT (tolerance 0).  The looser north-star bound (1e-5 relative) applies only to the
optional warp-shuffle epilogue (exact_sum = 0), tested with that tolerance."""
import ctypes

import numpy as np
import pytest

from util import CASES, bits_equal, float_table_equal, gold, oracle_cfg, product_config, rel_err, sha

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # north_star tolerance for fp32 power maps (non-exact epilogue only)


def _native():
    from lib import _native as nat
    return nat


def _load_case(case):
    config = product_config(case)
    nat = _native()
    nat.configure_from(config)
    nat.lib().bf_set_kernel_options(0, 1)
    return config, nat, nat.lib(), gold(case)


def _signals(case, g):
    if "signals" in g:
        return np.ascontiguousarray(g["signals"])
    from lib import synthetic
    s = np.ascontiguousarray(synthetic.plot_py_stimulus(256, 256))
    assert sha(s) == str(g["signals_sha"])
    return s


def _mimo(L, nat, name, sig, mics, D):
    img = np.full(D, np.nan, np.float32)
    getattr(L, name)(nat.ptr(sig), nat.ptr(img), nat.ptr(mics), len(mics))
    nat.check()
    return img


@pytest.mark.parametrize("case", list(CASES))
def test_device_delay_generator_bit_exact(case):
    config, nat, L, g = _load_case(case)
    from lib import directions
    delays = directions.calculate_delays()
    assert delays.dtype == np.float64 and tuple(g["grid"]) == delays.shape[:2]
    from oracle import directions_np as dn
    ref_delays = dn.calculate_delays(oracle_cfg(case))      # pinned to delays_sha by the CPU suite
    assert float_table_equal(delays, ref_delays)
    if not np.any(ref_delays == 0) or sha(delays) == str(g["delays_sha"]):
        assert sha(delays) == str(g["delays_sha"])
    whole, d32 = directions.whole_and_f32()
    assert sha(whole) == str(g["whole_sha"]) and float_table_equal(d32, np.float32(ref_delays))
    # load_coefficients_lerp's split, executed on the device
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    lw, lf = np.zeros(d32.size, np.int32), np.zeros(d32.size, np.float32)
    nat.check(L.bf_get_lerp_tables(nat.ptr(lw), nat.ptr(lf), d32.size))
    assert sha(lw) == str(g["lerp_whole_sha"]) and sha(lf) == str(g["lerp_weight_sha"])
    w2, _ = directions.calculate_coefficients() if case in ("ragged", "c1") else (whole, None)
    assert np.array_equal(np.asarray(w2), whole)


@pytest.mark.parametrize("simple", [0, 1])
@pytest.mark.parametrize("case", list(CASES))
def test_power_maps_vs_reference_golden(case, simple):
    config, nat, L, g = _load_case(case)
    if simple and case == "c3":
        pytest.skip("simple kernel at C3 size is covered by the tiled-vs-simple test")
    L.bf_set_kernel_options(simple, 1)
    from lib import directions
    sig, mics = _signals(case, g), nat.i32(g["mic_ids"])
    D, n = config.MAX_RES_X * config.MAX_RES_Y, len(mics)
    whole, d32 = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    nat.check()
    img = _mimo(L, nat, "mimo_pad", sig, mics, D)
    assert bits_equal(img, g["img_pad"]), np.nanmax(rel_err(img, g["img_pad"]))
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    img = _mimo(L, nat, "mimo_lerp", sig, mics, D)
    assert bits_equal(img, g["img_lerp"]), np.nanmax(rel_err(img, g["img_lerp"]))
    for i, d in enumerate(g["miso_dirs"]):
        out = np.full(config.N_SAMPLES, np.nan, np.float32)
        L.miso_pad(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n, int(d) * n)
        nat.check()
        assert bits_equal(out, g["miso_pad"][i])
        L.miso_lerp(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n, int(d) * n)
        nat.check()
        assert bits_equal(out, g["miso_lerp"][i])
    L.bf_set_kernel_options(0, 1)


@pytest.mark.parametrize("case", ["c1", "c3", "default", "ragged"])
def test_shared_sum_tolerance_mode(case):
    """exact_sum = 2 (opt-in): microphones whose delay is the same for all 8 directions of a group are
    summed once per group and added to the 8 direction sums at the end -- a different rounding order, so
    the bar is the north-star tolerance (1e-5 relative per pixel), not bit equality."""
    config, nat, L, g = _load_case(case)
    from lib import directions
    sig, mics = _signals(case, g), nat.i32(g["mic_ids"])
    D = config.MAX_RES_X * config.MAX_RES_Y
    whole, d32 = directions.whole_and_f32()
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    L.bf_set_kernel_options(0, 2)
    try:
        for name, ref in (("mimo_pad", g["img_pad"]), ("mimo_lerp", g["img_lerp"])):
            img = _mimo(L, nat, name, sig, mics, D)
            err = rel_err(img, ref)
            print("%s %s shared-sum: max per-pixel rel err %.3e (mean %.3e)" % (case, name, err.max(), err.mean()))
            assert err.max() <= REL_TOL, (name, float(err.max()))
    finally:
        L.bf_set_kernel_options(0, 1)


@pytest.mark.parametrize("case", ["c1", "ragged", "taps64"])
def test_fir_and_hybrid_vs_reference_golden(case):
    config, nat, L, g = _load_case(case)
    from lib import directions
    sig, mics = _signals(case, g), nat.i32(g["mic_ids"])
    D = config.MAX_RES_X * config.MAX_RES_Y
    taps = nat.f32(directions.compute_convolve_h())
    assert sha(taps) == str(g["taps_sha"])
    L.load_coefficients_convolve(nat.ptr(taps), taps.size)
    nat.check()
    assert bits_equal(_mimo(L, nat, "mimo_convolve_naive", sig, mics, D), g["img_fir_seq"])
    assert bits_equal(_mimo(L, nat, "mimo_convolve_vectorized", sig, mics, D), g["img_fir_lanes"])
    _, d32 = directions.whole_and_f32()
    L.load_coefficients_convolve_hybrid(nat.ptr(d32), d32.size)
    nat.check()
    hw = np.zeros(d32.size, np.int32)
    ht = np.zeros(d32.size * config.N_TAPS, np.float32)
    nat.check(L.bf_get_hybrid_tables(nat.ptr(hw), nat.ptr(ht), d32.size))
    assert sha(ht) == str(g["hybrid_taps_sha"])
    assert bits_equal(_mimo(L, nat, "mimo_convolve_hybrid", sig, mics, D), g["img_hybrid"])


def test_reference_wrappers_surface():
    """lib.tests.mimo_*_wrapper == the reference's own wrappers on the plot.py stimulus."""
    config, nat, L, g = _load_case("default")
    from lib import synthetic, tests as ltests
    sig = synthetic.plot_py_stimulus(config.N_MICROPHONES, config.N_SAMPLES)
    a = ltests.mimo_pad_wrapper(sig)
    assert a.shape == (57, 32) and a.dtype == np.float32 and bits_equal(a, g["wrapper_pad"])
    c = ltests.mimo_lerp_wrapper(sig)
    assert bits_equal(c, g["wrapper_lerp"])
    assert np.unravel_index(a.argmax(), a.shape) == (28, 14)
    out = ltests.pad_delay_wrapper(sig[0], np.ones(256, np.float32), 5)
    exp = np.ones(256, np.float32)
    exp[5:] += sig[0, :251]
    assert bits_equal(out, exp)


@pytest.mark.parametrize("n_samples", [64, 128, 256])
@pytest.mark.parametrize("n_mics,n_use", [(256, 256), (64, 37), (16, 1)])
def test_random_tables_vs_oracle(n_samples, n_mics, n_use):
    """Random delays anywhere in [0, N] (and a few out of range), non-power-of-two mic counts,
    shuffled adaptive arrays, ragged direction counts: tiled kernel == simple kernel == oracle."""
    from oracle import cpu
    nat = _native()
    L = nat.lib()
    X, Y = 13, 7                                    # D = 91: not a multiple of the 8-direction group
    nat.configure(n_mics, n_samples, 8, X, Y)
    D = X * Y
    rng = np.random.default_rng(n_samples * 1000 + n_mics + n_use)
    sig = rng.standard_normal((n_mics, n_samples)).astype(np.float32)
    mics = nat.i32(rng.permutation(n_mics)[:n_use])
    whole = rng.integers(0, n_samples + 1, (D, n_use)).astype(np.int32)
    whole[::5] = whole[::5, :1]                     # rows where all mics share one delay
    whole[3] = 0
    whole[4] = n_samples
    d32 = (rng.random((D, n_use)) * (n_samples - 1)).astype(np.float32)
    d32[7] = np.floor(d32[7])                       # exact integers: weight == 1
    ref_pad = cpu.mimo_pad(sig, mics, whole, D)
    ref_lerp = cpu.mimo_lerp(sig, mics, d32, D)
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    nat.check()
    for simple in (0, 1):
        L.bf_set_kernel_options(simple, 1)
        assert bits_equal(_mimo(L, nat, "mimo_pad", sig, mics, D), ref_pad), (simple, "pad")
        assert bits_equal(_mimo(L, nat, "mimo_lerp", sig, mics, D), ref_lerp), (simple, "lerp")
    # warp-shuffle epilogue: within the north-star tolerance
    L.bf_set_kernel_options(0, 0)
    got = _mimo(L, nat, "mimo_pad", sig, mics, D)
    assert np.all(np.abs(got - ref_pad) <= REL_TOL * np.abs(ref_pad) + 1e-30)
    got = _mimo(L, nat, "mimo_lerp", sig, mics, D)
    assert np.all(np.abs(got - ref_lerp) <= REL_TOL * np.abs(ref_lerp) + 1e-30)
    L.bf_set_kernel_options(0, 1)
    off = 11 * n_use
    for name, fn, tab in (("miso_pad", cpu.miso_pad, whole), ("miso_lerp", cpu.miso_lerp, d32)):
        out = np.zeros(n_samples, np.float32)
        getattr(L, name)(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n_use, off)
        nat.check()
        assert bits_equal(out, fn(sig, mics, tab, off)), name
    by_mic = rng.integers(0, n_samples // 2, n_mics).astype(np.int32)
    L.load_coefficients_pad2(nat.ptr(by_mic), by_mic.size)
    out = np.zeros(n_samples, np.float32)
    L.miso_pad2(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n_use, 0)
    nat.check()
    assert bits_equal(out, cpu.miso_pad2(sig, mics, by_mic))


def test_single_row_delay_entry_points():
    from oracle import cpu
    nat = _native()
    L = nat.lib()
    N, T = 256, 8
    nat.configure(4, N, T, 2, 2)
    rng = np.random.default_rng(9)
    row = rng.standard_normal(N).astype(np.float32)
    base = rng.standard_normal(N).astype(np.float32)
    h = rng.standard_normal(T).astype(np.float32)
    O = cpu.lib()
    def orc(fn, *a):
        out = base.copy()
        fn(*[x if not isinstance(x, str) else nat.ptr(out) for x in a])
        return out
    out = base.copy(); L.pad_delay(nat.ptr(row), nat.ptr(out), 17); nat.check()
    assert bits_equal(out, orc(O.orc_pad_delay, nat.ptr(row), "out", 17, N))
    out = base.copy(); L.lerp_delay(nat.ptr(row), nat.ptr(out), ctypes.c_float(0.375), 9); nat.check()
    assert bits_equal(out, orc(O.orc_lerp_delay, nat.ptr(row), "out", ctypes.c_float(0.375), 9, N))
    out = base.copy(); L.convolve_delay_naive_add(nat.ptr(row), nat.ptr(h), nat.ptr(out)); nat.check()
    assert bits_equal(out, orc(O.orc_fir_delay_seq, nat.ptr(row), nat.ptr(h), "out", N, T))
    out = base.copy(); L.convolve_delay_naive(nat.ptr(row), nat.ptr(out), nat.ptr(h)); nat.check()
    assert bits_equal(out, orc(O.orc_fir_delay_seq, nat.ptr(row), nat.ptr(h), "out", N, T))
    out = base.copy(); L.convolve_delay_vectorized_add(nat.ptr(row), nat.ptr(h), nat.ptr(out)); nat.check()
    assert bits_equal(out, orc(O.orc_fir_delay_lanes, nat.ptr(row), nat.ptr(h), "out", N, T))
    out = base.copy(); L.convolve_hybrid_delay_add(nat.ptr(row), nat.ptr(h), 21, nat.ptr(out)); nat.check()
    assert bits_equal(out, orc(O.orc_hybrid_delay, nat.ptr(row), nat.ptr(h), 21, "out", N, T))


def test_misuse_is_reported_not_undefined():
    nat = _native()
    L = nat.lib()
    nat.configure(16, 256, 8, 4, 4)
    L.unload_coefficients_pad()
    sig = np.zeros((16, 256), np.float32)
    img = np.zeros(16, np.float32)
    mics = np.arange(16, dtype=np.int32)
    L.mimo_pad(nat.ptr(sig), nat.ptr(img), nat.ptr(mics), 16)
    assert L.bf_last_status() == 2                      # BF_ERR_NOT_LOADED
    small = np.zeros(10, np.int32)
    L.load_coefficients_pad(nat.ptr(small), 10)
    L.mimo_pad(nat.ptr(sig), nat.ptr(img), nat.ptr(mics), 16)
    assert L.bf_last_status() == 2
    bad = np.array([0, 99], np.int32)
    L.mimo_pad(nat.ptr(sig), nat.ptr(img), nat.ptr(bad), 2)
    assert L.bf_last_status() == 4                      # BF_ERR_ARG
    L.pad_mimo(nat.ptr(img), nat.ptr(mics), 16)
    assert L.bf_last_status() != 0 and b"data source" in L.bf_last_error()


@pytest.mark.parametrize("case", ["ragged", "taps64"])
def test_miso_fir_and_hybrid_vs_oracle(case):
    """miso_convolve_naive / miso_convolve_vectorized / miso_convolve_hybrid:
    single-direction FIR beams, both accumulation orders, 8 and 64 taps; offsets are in floats
    (d*n*T) for the FIR table and in entries (d*n) for the hybrid one, as in the reference."""
    from oracle import cpu
    config, nat, L, g = _load_case(case)
    from lib import directions
    sig, mics = _signals(case, g), nat.i32(g["mic_ids"])
    n, T, N = len(mics), config.N_TAPS, config.N_SAMPLES
    taps = nat.f32(directions.compute_convolve_h())
    L.load_coefficients_convolve(nat.ptr(taps), taps.size)
    _, d32 = directions.whole_and_f32()
    L.load_coefficients_convolve_hybrid(nat.ptr(d32), d32.size)
    nat.check()
    for d in (0, 7, config.MAX_RES_X * config.MAX_RES_Y - 1):
        out = np.full(N, np.nan, np.float32)
        L.miso_convolve_naive(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n, d * n * T)
        nat.check()
        assert bits_equal(out, cpu.miso_fir(sig, mics, taps, d * n * T, T, 0)), ("naive", d)
        L.miso_convolve_vectorized(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n, d * n * T)
        nat.check()
        assert bits_equal(out, cpu.miso_fir(sig, mics, taps, d * n * T, T, 1)), ("vectorized", d)
        L.miso_convolve_hybrid(nat.ptr(sig), nat.ptr(out), nat.ptr(mics), n, d * n)
        nat.check()
        assert bits_equal(out, cpu.miso_hybrid(sig, mics, d32, d * n, T)), ("hybrid", d)


def test_api_h_wrappers_over_a_data_source():
    """api.h:11-20 -- pad_mimo / lerp_mimo / convolve_mimo_* / mimo_truncated / miso_steer_listen fetch
    the buffer from the registered source (get_data() in the reference) and then run the kernels:
    same result as the explicit-buffer calls; load_coefficients2 feeds mimo_truncated (api.c:1004-1087)."""
    config, nat, L, g = _load_case("ragged")
    from lib import beamformer, directions
    sig, mics = _signals("ragged", g), nat.i32(g["mic_ids"])
    n, D, N = len(mics), config.MAX_RES_X * config.MAX_RES_Y, config.N_SAMPLES
    whole, d32 = directions.whole_and_f32()
    taps = nat.f32(directions.compute_convolve_h())
    L.load_coefficients_pad(nat.ptr(whole), whole.size)
    L.load_coefficients2(nat.ptr(whole), whole.size)
    L.load_coefficients_lerp(nat.ptr(d32), d32.size)
    L.load_coefficients_convolve(nat.ptr(taps), taps.size)
    nat.check()
    beamformer.connect(False, verbose=False, source=beamformer.ArraySource(sig))
    try:
        buf = np.empty((config.N_MICROPHONES, N), np.float32)
        beamformer.receive(buf)
        assert bits_equal(buf, sig)
        for wrapper, direct in (("pad_mimo", "mimo_pad"), ("lerp_mimo", "mimo_lerp"), ("mimo_truncated", "mimo_pad"),
                                ("convolve_mimo_naive", "mimo_convolve_naive"),
                                ("convolve_mimo_vectorized", "mimo_convolve_vectorized")):
            a = np.full(D, np.nan, np.float32)
            getattr(L, wrapper)(nat.ptr(a), nat.ptr(mics), n)
            nat.check()
            assert bits_equal(a, _mimo(L, nat, direct, sig, mics, D)), wrapper
        assert bits_equal(_mimo(L, nat, "mimo_pad", sig, mics, D), g["img_pad"])
        out = np.full(N, np.nan, np.float32)
        L.miso_steer_listen(nat.ptr(out), nat.ptr(mics), n, 38 * n)
        nat.check()
        assert bits_equal(out, g["miso_pad"][1])
    finally:
        beamformer.disconnect()


def test_degenerate_inputs():
    """Edge cases of the calling surface: a single direction, a single microphone, delays at and
    beyond the block length, the largest supported block (simple kernel, N = 1024)."""
    from oracle import cpu
    nat = _native()
    L = nat.lib()
    rng = np.random.default_rng(12)
    for N, X, Y, M, n in ((256, 1, 1, 4, 4), (1024, 3, 2, 8, 5), (32, 5, 1, 3, 1)):
        nat.configure(M, N, 8, X, Y)
        D = X * Y
        sig = rng.standard_normal((M, N)).astype(np.float32)
        mics = nat.i32(rng.permutation(M)[:n])
        whole = rng.integers(0, N + 5, (D, n)).astype(np.int32)        # some delays beyond the block
        d32 = (rng.random((D, n)) * (N + 3)).astype(np.float32)
        L.load_coefficients_pad(nat.ptr(whole), whole.size)
        L.load_coefficients_lerp(nat.ptr(d32), d32.size)
        nat.check()
        assert bits_equal(_mimo(L, nat, "mimo_pad", sig, mics, D), cpu.mimo_pad(sig, mics, whole, D)), (N, "pad")
        assert bits_equal(_mimo(L, nat, "mimo_lerp", sig, mics, D), cpu.mimo_lerp(sig, mics, d32, D)), (N, "lerp")
    nat.configure(256, 256, 8, 57, 32)


@pytest.mark.parametrize("taps", [8, 16, 32])
@pytest.mark.parametrize("n_mics,n_use", [(256, 256), (64, 37)])
def test_fir_tiled_random_vs_oracle(taps, n_mics, n_use):
    """Tiled FIR kernel (csrc/das_fir.cu) == one-thread-per-sample kernel == oracle, bit for bit: random
    taps, a direction count that is not a multiple of the 8-direction group, a non-power-of-two number
    of shuffled microphones, fused (T <= 16) and unfused (T = 32) chains, both accumulation orders, a
    batch of frames, a direction slice written through strides, and the shuffle-tree epilogue."""
    import torch
    from oracle import cpu
    nat = _native()
    L = nat.lib()
    X, Y, N, F = 13, 7, 256, 3
    D = X * Y
    nat.configure(n_mics, N, taps, X, Y)
    rng = np.random.default_rng(taps * 100 + n_use)
    sig = rng.standard_normal((F, n_mics, N)).astype(np.float32)
    mics = nat.i32(rng.permutation(n_mics)[:n_use])
    h = (rng.standard_normal((D, n_use, taps)) / taps).astype(np.float32)
    h[5] = 0.0
    h[6, :, 1:] = 0.0
    L.load_coefficients_convolve(nat.ptr(h), h.size)
    nat.check()
    d_sig, d_mics = torch.from_numpy(sig).cuda(), torch.from_numpy(mics).cuda()
    for algo, lanes in ((nat.ALGO_FIR_SEQ, 0), (nat.ALGO_FIR_LANES, 1)):
        ref = np.stack([cpu.mimo_fir(sig[f], mics, h, D, taps, lanes) for f in range(F)])
        for simple in (0, 1):
            L.bf_set_kernel_options(simple, 1)
            d_img = torch.full((F, D), float("nan"), device="cuda")
            nat.check(L.bf_mimo_dev(algo, d_sig.data_ptr(), d_img.data_ptr(), F, d_mics.data_ptr(), n_use, 0, D, None))
            torch.cuda.synchronize()
            assert bits_equal(d_img.cpu().numpy(), ref), (algo, simple)
        # direction slice [19, 19+40) written direction-major with origin 19
        L.bf_set_kernel_options(0, 1)
        d_sl = torch.full((40, F), float("nan"), device="cuda")
        nat.check(L.bf_mimo_dev_ex(algo, d_sig.data_ptr(), d_sl.data_ptr(), F, d_mics.data_ptr(), n_use, 19, 40,
                                   1, F, 19, None))
        torch.cuda.synchronize()
        assert bits_equal(np.ascontiguousarray(d_sl.cpu().numpy().T), ref[:, 19:59]), algo
        # shuffle-tree epilogue: tolerance only
        L.bf_set_kernel_options(0, 0)
        d_img = torch.zeros((F, D), device="cuda")
        nat.check(L.bf_mimo_dev(algo, d_sig.data_ptr(), d_img.data_ptr(), F, d_mics.data_ptr(), n_use, 0, D, None))
        torch.cuda.synchronize()
        got = d_img.cpu().numpy()
        assert np.all(np.abs(got - ref) <= REL_TOL * np.abs(ref) + 1e-30)
        L.bf_set_kernel_options(0, 1)
    # the host-pointer drop-in names go through the same kernels
    assert bits_equal(_mimo(L, nat, "mimo_convolve_naive", sig[0], mics, D), cpu.mimo_fir(sig[0], mics, h, D, taps, 0))
