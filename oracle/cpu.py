"""ctypes front-end of oracle/oracle.c (the CPU restatement).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/oracle.c.  Importing this
module compiles oracle.c with gcc into oracle/_build/liboracle.so when missing
or stale.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle.c")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "liboracle.so")

# -ffp-contract=off: every fused multiply-add is written explicitly (fmaf);
# -mfma so that fmaf() is one instruction; no -march=native (the .so may be
# built on one host and run on another).
CFLAGS = ["-O3", "-mavx2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
          "-fvisibility=hidden"]


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", *CFLAGS, SRC, "-o", LIB, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_time_mimo.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def mimo_pad(signals, mic_ids, whole, D):
    signals, mic_ids, whole = _f32(signals), _i32(mic_ids), _i32(whole).ravel()
    N = signals.shape[1]
    img = np.zeros(D, np.float32)
    lib().orc_mimo_pad(_p(signals), _p(img), _p(mic_ids), len(mic_ids), _p(whole), D, N)
    return img


def miso_pad(signals, mic_ids, whole, offset):
    signals, mic_ids, whole = _f32(signals), _i32(mic_ids), _i32(whole).ravel()
    N = signals.shape[1]
    out = np.zeros(N, np.float32)
    lib().orc_miso_pad(_p(signals), _p(out), _p(mic_ids), len(mic_ids), _p(whole), int(offset), N)
    return out


def miso_pad2(signals, mic_ids, whole_by_mic):
    signals, mic_ids, w = _f32(signals), _i32(mic_ids), _i32(whole_by_mic).ravel()
    N = signals.shape[1]
    out = np.zeros(N, np.float32)
    lib().orc_miso_pad2(_p(signals), _p(out), _p(mic_ids), len(mic_ids), _p(w), N)
    return out


def split_lerp(delays_f32):
    d = _f32(delays_f32).ravel()
    whole = np.zeros(d.size, np.int32)
    weight = np.zeros(d.size, np.float32)
    lib().orc_split_lerp(_p(d), d.size, _p(whole), _p(weight))
    return whole, weight


def mimo_lerp(signals, mic_ids, delays_f32, D):
    signals, mic_ids = _f32(signals), _i32(mic_ids)
    whole, weight = split_lerp(delays_f32)
    N = signals.shape[1]
    img = np.zeros(D, np.float32)
    lib().orc_mimo_lerp(_p(signals), _p(img), _p(mic_ids), len(mic_ids), _p(whole), _p(weight), D, N)
    return img


def miso_lerp(signals, mic_ids, delays_f32, offset):
    signals, mic_ids = _f32(signals), _i32(mic_ids)
    whole, weight = split_lerp(delays_f32)
    N = signals.shape[1]
    out = np.zeros(N, np.float32)
    lib().orc_miso_lerp(_p(signals), _p(out), _p(mic_ids), len(mic_ids), _p(whole), _p(weight),
                        int(offset), N)
    return out


def mimo_fir(signals, mic_ids, taps, D, T, lanes):
    signals, mic_ids, taps = _f32(signals), _i32(mic_ids), _f32(taps).ravel()
    N = signals.shape[1]
    img = np.zeros(D, np.float32)
    lib().orc_mimo_fir(_p(signals), _p(img), _p(mic_ids), len(mic_ids), _p(taps), D, N, T, int(lanes))
    return img


def miso_fir(signals, mic_ids, taps, offset, T, lanes):
    signals, mic_ids, taps = _f32(signals), _i32(mic_ids), _f32(taps).ravel()
    N = signals.shape[1]
    out = np.zeros(N, np.float32)
    lib().orc_miso_fir(_p(signals), _p(out), _p(mic_ids), len(mic_ids), _p(taps), int(offset), N, T,
                       int(lanes))
    return out


def split_hybrid(delays_f32, T):
    d = _f32(delays_f32).ravel()
    whole = np.zeros(d.size, np.int32)
    taps = np.zeros(d.size * T, np.float32)
    lib().orc_split_hybrid(_p(d), d.size, _p(whole), _p(taps), T)
    return whole, taps


def mimo_hybrid(signals, mic_ids, delays_f32, D, T):
    signals, mic_ids = _f32(signals), _i32(mic_ids)
    whole, taps = split_hybrid(delays_f32, T)
    N = signals.shape[1]
    img = np.zeros(D, np.float32)
    lib().orc_mimo_hybrid(_p(signals), _p(img), _p(mic_ids), len(mic_ids), _p(whole), _p(taps), D, N, T)
    return img


def miso_scale(out, n, gain):
    out = _f32(out).copy()
    lib().orc_miso_scale(_p(out), out.size, int(n), ctypes.c_float(gain))
    return out


def ingest(stream_i32, n_arrays, rows=8, cols=8, norm=16777216.0, quirk=True):
    s = _i32(stream_i32)
    N, M = s.shape
    out = np.zeros((M, N), np.float32)
    lib().orc_ingest(_p(s), _p(out), N, M, int(n_arrays), rows, cols, ctypes.c_double(norm), int(quirk))
    return out


def time_mimo(kind, signals, mic_ids, whole, weight, Dsub, reps):
    """Seconds for `reps` passes over the first Dsub directions (bench cpu_baseline)."""
    signals, mic_ids, whole = _f32(signals), _i32(mic_ids), _i32(whole).ravel()
    weight = _f32(weight).ravel() if weight is not None else np.zeros(1, np.float32)
    N = signals.shape[1]
    img = np.zeros(Dsub, np.float32)
    return lib().orc_time_mimo(int(kind), _p(signals), _p(img), _p(mic_ids), len(mic_ids), _p(whole),
                               _p(weight), int(Dsub), N, int(reps))


def miso_hybrid(signals, mic_ids, delays_f32, offset, T):
    signals, mic_ids = _f32(signals), _i32(mic_ids)
    whole, taps = split_hybrid(delays_f32, T)
    N = signals.shape[1]
    out = np.zeros(N, np.float32)
    lib().orc_miso_hybrid(_p(signals), _p(out), _p(mic_ids), len(mic_ids), _p(whole), _p(taps), int(offset), N, T)
    return out
