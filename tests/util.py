"""Shared helpers for the test-suite (imports the oracle: tests are allowed to)."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# config.json overrides of each golden case (mirrors oracle/build_ref.py CONFIGS)
CASES = {
    "default": dict(),
    "c1": dict(N_MICROPHONES=64, MAX_RES_X=20, MAX_RES_Y=20, GEOMETRY_N_MICS=64, GEOMETRY_N_ARRAYS=1),
    "c3": dict(MAX_RES_X=180, MAX_RES_Y=180),
    "ragged": dict(MAX_RES_X=11, MAX_RES_Y=7, SKIP_N_MICS=3),
    "taps64": dict(MAX_RES_X=9, MAX_RES_Y=5, N_TAPS=64),
}
BASE = dict(N_MICROPHONES=256, N_SAMPLES=256, N_TAPS=8, MAX_RES_X=57, MAX_RES_Y=32, SKIP_N_MICS=1,
            GEOMETRY_N_MICS=256, GEOMETRY_N_ARRAYS=4)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def oracle_cfg(case):
    """oracle/directions_np.py config dict for a golden case."""
    from oracle import directions_np as dn
    kw = dict(CASES[case])
    if "GEOMETRY_N_MICS" in kw:
        kw["n_mics_geom"] = kw.pop("GEOMETRY_N_MICS")
        kw["n_arrays_geom"] = kw.pop("GEOMETRY_N_ARRAYS")
    return dn.cfg_with(**kw)


def product_config(case):
    """Point the product's interface.config at a golden case; returns the module."""
    from interface import config
    kw = dict(BASE)
    kw.update(CASES[case])
    return config.reload(**kw)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def float_table_equal(a, b):
    """Bitwise equality of float tables, except that +0.0 and -0.0 are interchangeable.  The only
    place they can differ is the on-axis direction of an odd x odd grid, where every raw delay is
    +-0 and `x - min(x)` inherits the sign from whichever zero NumPy's SIMD min-reduction happened
    to return on the build host; the integer table, the lerp split and every power map are
    unaffected (documented in DESIGN.md)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    u = np.uint64 if a.dtype == np.float64 else np.uint32
    same = a.view(u) == b.view(u)
    return bool(np.all(same | ((a == 0) & (b == 0))))
