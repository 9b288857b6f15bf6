"""realtime_scripts/beam_forming_algorithm.py -- main(signal) on the GPU.

signal: float32 (N_SAMPLES, N_MICROPHONES) -- the transposed sample buffer the reference passes
(`main(data.T)`, PC/application/camera.py:72).  Returns the float64 (MAX_RES_X, MAX_RES_Y) heat
map: sum over bins of |sum over mics X * exp(j*phase)|^2, divided by its maximum, or all zeros
when the maximum is below threshold_heatmap (beam_forming_algorithm.py:50-63)."""
import numpy as np

import realtime_scripts.calc_phase_shift_cartesian as calc_phase_shift_cartesian
import realtime_scripts.config as config
from lib import _native

n_samples = calc_phase_shift_cartesian.N
f = calc_phase_shift_cartesian.f
active_mics = calc_phase_shift_cartesian.active_mics
x_scan = calc_phase_shift_cartesian.x_scan
y_scan = calc_phase_shift_cartesian.y_scan
x_res = config.MAX_RES_X
y_res = config.MAX_RES_Y

threshold_heatmap = 0.2
threshold_freq_lower_idx = calc_phase_shift_cartesian.threshold_freq_lower_idx
threshold_freq_upper_idx = calc_phase_shift_cartesian.threshold_freq_upper_idx

_installed = False


def _ensure():
    global _installed
    if not _installed:
        calc_phase_shift_cartesian.install()
        _installed = True


def power(signal):
    """Un-normalised sum over bins of the steered power, float32 (x_res, y_res)."""
    _ensure()
    sig = np.ascontiguousarray(np.asarray(signal, dtype=np.float32).T)        # (M, N)
    assert sig.shape == (config.N_MICROPHONES, n_samples), "Arrays do not match shape"
    out = np.zeros((x_res, y_res), dtype=np.float32)
    _native.check(_native.lib().bf_fd_das(_native.ptr(sig), _native.ptr(out), 1, threshold_heatmap, 0))
    return out


def main(signal):
    _ensure()
    sig = np.ascontiguousarray(np.asarray(signal, dtype=np.float32).T)        # (M, N)
    assert sig.shape == (config.N_MICROPHONES, n_samples), "Arrays do not match shape"
    out = np.zeros((x_res, y_res), dtype=np.float32)
    _native.check(_native.lib().bf_fd_das(_native.ptr(sig), _native.ptr(out), 1, threshold_heatmap, 1))
    return out.astype(np.float64)
