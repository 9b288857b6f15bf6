#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REAL reference build (oracle/_ref).

Run in the build container only (needs oracle/_ref from oracle/build_ref.py and,
for the frequency-domain fixture, /root/reference itself, whose NumPy FD module
is imported in place).  The fixtures are small, committed, and are what pins the
oracle restatement and the CUDA path on the GPU box, where neither
/root/reference nor a compiler run against it exists.

    python oracle/build_ref.py && python oracle/gen_golden.py
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200"))

from oracle import ref  # noqa: E402
from lib import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("BF_REFERENCE_ROOT", "/root/reference")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %s (%.1f KiB)" % (path, os.path.getsize(path) / 1024))


def td_case(cfg, signals, store_table, with_fir, miso_dirs):
    R = ref.RefC(cfg)
    dm = ref.directions(cfg)
    mics, n = dm.active_microphones()
    mics = np.asarray(mics, np.int32)
    delays = dm.calculate_delays()                       # float64 (X, Y, n)
    whole = delays.astype(int).astype(np.int32)          # == calculate_coefficients()[0]
    d32 = np.float32(delays)
    out = dict(mic_ids=mics, signals=signals,
               delays_sha=np.array(sha(delays)), whole_sha=np.array(sha(whole)),
               d32_sha=np.array(sha(d32)), grid=np.array(delays.shape[:2]),
               r_prime=dm.calc_r_prime(float(np.float32(0.02))))
    if store_table:
        out["delays"] = delays
    lw, lf = R.lerp_tables(d32)
    out["lerp_whole_sha"] = np.array(sha(lw))
    out["lerp_weight_sha"] = np.array(sha(lf))
    out["img_pad"] = R.mimo_pad(signals, mics, whole)
    out["img_lerp"] = R.mimo_lerp(signals, mics, d32)
    out["miso_dirs"] = np.array(miso_dirs, np.int32)
    out["miso_pad"] = np.stack([R.miso_pad(signals, mics, whole, d * n) for d in miso_dirs])
    out["miso_lerp"] = np.stack([R.miso_lerp(signals, mics, d32, d * n) for d in miso_dirs])
    if with_fir:
        taps = dm.compute_convolve_h()                   # float32 (X, Y, n, T)
        out["taps_sha"] = np.array(sha(taps))
        out["img_fir_seq"] = R.mimo_fir(signals, mics, taps, lanes=0)
        out["img_fir_lanes"] = R.mimo_fir(signals, mics, taps, lanes=1)
        hw, ht = R.hybrid_tables(d32)
        out["hybrid_taps_sha"] = np.array(sha(ht))
        out["img_hybrid"] = R.mimo_hybrid(signals, mics, d32)
    return out, delays, mics


def gen_c1():
    dm = ref.directions("c1")
    mics, n = dm.active_microphones()
    delays = dm.calculate_delays().reshape(-1, n)
    w = synthetic.C1
    sig = synthetic.point_sources(delays, mics, w["n_mics_total"], 256, 48828.0, w["sources"],
                                  w["noise"], w["seed"])
    out, _, _ = td_case("c1", sig, store_table=True, with_fir=True, miso_dirs=[0, 14 * 20 + 6, 399])
    save("c1", **out)


def gen_ragged():
    R = ref.RefC("ragged")
    rng = np.random.default_rng(77)
    sig = rng.standard_normal((R.M, R.N)).astype(np.float32)
    out, _, _ = td_case("ragged", sig, store_table=True, with_fir=True, miso_dirs=[0, 38, 76])
    save("ragged", **out)


def gen_taps64():
    R = ref.RefC("taps64")
    rng = np.random.default_rng(78)
    sig = rng.standard_normal((R.M, R.N)).astype(np.float32)
    out, _, _ = td_case("taps64", sig, store_table=False, with_fir=True, miso_dirs=[0, 44])
    save("taps64", **out)


def gen_default():
    """Stock config + the reference's own stimulus, through the reference's own
    Python wrappers (benchmark.pyx: mimo_pad_wrapper / mimo_lerp_wrapper)."""
    R = ref.RefC("default")
    sig = synthetic.plot_py_stimulus(R.M, R.N)
    out, delays, mics = td_case("default", sig, store_table=False, with_fir=False,
                                miso_dirs=[0, 28 * 32 + 14, 1823])
    del out["signals"]                                   # regenerated from plot_py_stimulus()
    out["signals_sha"] = np.array(sha(sig))
    out["whole_i8"] = delays.astype(int).astype(np.int8)  # max delay 47 -> fits int8
    out["wrapper_pad"], out["wrapper_lerp"] = ref.run_ref_wrappers(sig)  # ~25 s (python tap loop)
    assert np.array_equal(out["wrapper_pad"].ravel(), out["img_pad"])
    assert np.array_equal(out["wrapper_lerp"].ravel(), out["img_lerp"])
    print("plot.py stimulus: pad max %.7f at %s, lerp max %.7f at %s" % (
        out["wrapper_pad"].max(), np.unravel_index(out["wrapper_pad"].argmax(), (57, 32)),
        out["wrapper_lerp"].max(), np.unravel_index(out["wrapper_lerp"].argmax(), (57, 32))))
    save("default", **out)


def gen_c3():
    dm = ref.directions("c3")
    mics, n = dm.active_microphones()
    delays = dm.calculate_delays().reshape(-1, n)
    w = synthetic.C3
    sig = synthetic.point_sources(delays, mics, w["n_mics_total"], 256, 48828.0, w["sources"],
                                  w["noise"], w["seed"])
    out, _, _ = td_case("c3", sig, store_table=False, with_fir=False,
                        miso_dirs=[0, 40 * 180 + 100, 32399])
    save("c3", **out)


def gen_fd():
    """Frequency-domain DAS: the reference's pure-NumPy module imported in place
    (PC/application/realtime_scripts), Matplotlib stubbed (only used for plots)."""
    app = os.path.join(REF, "PC", "application")
    if not os.path.isdir(app):
        print("reference not present: skipping FD fixture")
        return
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].rc = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    cwd = os.getcwd()
    os.chdir(HERE)                      # a cwd WITHOUT unused_mics.npy, like the live apps
    sys.path.insert(0, app)
    try:
        import realtime_scripts.beam_forming_algorithm as bfa
        import realtime_scripts.calc_phase_shift_cartesian as cps
    finally:
        os.chdir(cwd)
    rng = np.random.default_rng(1237)
    # one broadband source steered at grid cell (9, 4) + noise, as float32 (N, M) = data.T
    F, M = cps.phase_shift.shape[0], cps.phase_shift.shape[1]
    N = cps.N
    t = np.arange(N)
    x_i, y_i = cps.x_i[0, :, 0, 0], cps.y_i[0, :, 0, 0]
    xs, ys = cps.x_scan[0, 0, 9, 0], cps.y_scan[0, 0, 0, 4]
    rs = np.sqrt(xs ** 2 + ys ** 2 + 1.0)
    tau = (xs * x_i + ys * y_i) / rs / cps.c * cps.fs            # samples
    sig = np.zeros((N, M))
    for f0, amp in ((3000.0, 0.2), (7000.0, 0.1)):
        sig += amp * np.sin(2 * np.pi * f0 * (t[:, None] + tau[None, :]) / cps.fs)
    sig += rng.normal(0, 0.01, sig.shape)
    sig = sig.astype(np.float32)
    heat = bfa.main(sig)
    quiet = bfa.main((sig * 1e-4).astype(np.float32))            # below threshold -> zeros
    save("fd_das",
         signal=sig, heatmap=heat, heatmap_quiet=quiet,
         r_prime_all=np.stack([x_i, y_i]), x_scan=cps.x_scan.ravel(), y_scan=cps.y_scan.ravel(),
         f=cps.f.ravel(), lo=np.array(cps.threshold_freq_lower_idx),
         hi=np.array(cps.threshold_freq_upper_idx), c=np.array(float(cps.c)),
         fs=np.array(int(cps.fs)), phase_sha=np.array(sha(cps.phase_shift)),
         fft_power=(np.abs(np.sum(bfa.frequency_phase_shift(sig, bfa.phase_shift), axis=1)) ** 2).sum(0))
    print("fd: phase_shift", cps.phase_shift.shape, "heat max at",
          np.unravel_index(heat.argmax(), heat.shape))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["c1", "ragged", "taps64", "default", "c3", "fd"]
    for w in which:
        globals()["gen_" + w]()
