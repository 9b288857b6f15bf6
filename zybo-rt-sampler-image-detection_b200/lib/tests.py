"""`lib.tests` -- the offline single-call wrappers of the reference
(PC/src/benchmark.pyx, built there as the extension module `tests`).

Same names and behaviour: each wrapper takes a float32 (N_MICROPHONES, N_SAMPLES)
array, builds the coefficient table for the configured geometry, loads it, runs
the power-map kernel and returns the float32 (MAX_RES_X, MAX_RES_Y) image
(flat direction index d = image.ravel()[d]).  All compute is on
the GPU through libbf_b200.so.
"""
import numpy as np

from interface import config
from . import _native
from .directions import (active_microphones, calculate_coefficients, calculate_delays,
                         compute_convolve_h, whole_and_f32)

DTYPE_arr = np.float32


def _signals(signals):
    s = _native.f32(signals)
    if s.shape != (config.N_MICROPHONES, config.N_SAMPLES):
        raise AssertionError("Arrays do not match shape")      # cf. main.pyx:154
    return s


def _mics():
    mics, n = active_microphones()
    return _native.i32(mics), int(n)


def _run(fn_name, signals):
    L = _native.lib()
    s = _signals(signals)
    mics, n = _mics()
    image = np.zeros((config.MAX_RES_X, config.MAX_RES_Y), dtype=DTYPE_arr)
    getattr(L, fn_name)(_native.ptr(s), _native.ptr(image), _native.ptr(mics), n)
    _native.check()
    return image


def pad_coefficients_load(whole_samples, n):
    """benchmark.pyx:58-72"""
    _native.configure_from(config)
    w = _native.i32(whole_samples)
    _native.lib().load_coefficients_pad(_native.ptr(w), int(w.size))
    _native.check()


def pad_delay_wrapper(signal, out, pos_pad) -> np.ndarray:
    """benchmark.pyx:74-82"""
    _native.configure_from(config)
    s, o = _native.f32(signal), _native.f32(out).copy()
    _native.lib().pad_delay(_native.ptr(s), _native.ptr(o), int(pos_pad))
    _native.check()
    return o


def mimo_pad_wrapper(signals):
    """benchmark.pyx:84-112 (calculate_coefficients -> load_coefficients_pad -> mimo_pad).
    The discarded FIR-tap half of calculate_coefficients() is not computed."""
    _native.configure_from(config)
    whole, _ = whole_and_f32()
    w = _native.i32(whole)
    _native.lib().load_coefficients_pad(_native.ptr(w), int(w.size))
    _native.check()
    return _run("mimo_pad", signals)


def convolve_coefficients_load(h):
    """benchmark.pyx:115-120"""
    _native.configure_from(config)
    t = _native.f32(h)
    _native.lib().load_coefficients_convolve(_native.ptr(t), int(t.size))
    _native.check()


def mimo_convolve_wrapper(signals):
    """benchmark.pyx:122-139 (compute_convolve_h -> load -> mimo_convolve_vectorized)"""
    convolve_coefficients_load(compute_convolve_h())
    return _run("mimo_convolve_vectorized", signals)


def mimo_lerp_wrapper(signals):
    """benchmark.pyx:141-162 (float32(calculate_delays()) -> load_coefficients_lerp ->
    mimo_lerp -> unload)"""
    _native.configure_from(config)
    _, d32 = whole_and_f32()
    _native.lib().load_coefficients_lerp(_native.ptr(d32), int(d32.size))
    _native.check()
    img = _run("mimo_lerp", signals)
    _native.lib().unload_coefficients_lerp()
    return img


def mimo_hybrid_convolve_wrapper(signals):
    """benchmark.pyx:164-186"""
    _native.configure_from(config)
    _, d32 = whole_and_f32()
    _native.lib().load_coefficients_convolve_hybrid(_native.ptr(d32), int(d32.size))
    _native.check()
    print("Could load")
    img = _run("mimo_convolve_hybrid", signals)
    print("Could convolve")
    return img


__all__ = ["pad_coefficients_load", "pad_delay_wrapper", "mimo_pad_wrapper",
           "convolve_coefficients_load", "mimo_convolve_wrapper", "mimo_lerp_wrapper",
           "mimo_hybrid_convolve_wrapper", "calculate_coefficients", "calculate_delays"]
