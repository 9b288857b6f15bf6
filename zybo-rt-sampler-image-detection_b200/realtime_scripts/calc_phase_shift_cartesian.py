"""realtime_scripts/calc_phase_shift_cartesian.py:8-50 without the table: the scan axes, the
frequency axis and the band limits are computed here exactly as the reference does (O(X+Y+N)
NumPy), and handed to the device with the microphone geometry (bf_fd_setup).  The reference's
`phase_shift` array (complex128 [F][M][X][Y]) is generated on the fly inside the steering kernel."""
import numpy as np

import realtime_scripts.active_microphones as am
import realtime_scripts.calc_r_prime as calc_r_prime
import realtime_scripts.config as config
from lib import _native

c = config.PROPAGATION_SPEED
fs = int(config.fs)
N = config.N_SAMPLES
d = config.ELEMENT_DISTANCE
theta_max = config.VIEW_ANGLE / 2
active_mics = am.active_microphones()

r_prime_all, r_prime = calc_r_prime.calc_r_prime(d)
x_i = r_prime_all[0, :]
y_i = r_prime_all[1, :]

x_scan_max = config.Z * np.tan(np.deg2rad(theta_max))
x_scan_min = -x_scan_max
y_scan_max = x_scan_max / config.ASPECT_RATIO
y_scan_min = -y_scan_max
x_scan = np.linspace(x_scan_min, x_scan_max, config.MAX_RES_X)
y_scan = np.linspace(y_scan_min, y_scan_max, config.MAX_RES_Y)

f = np.linspace(0, int(fs / 2), int(N / 2) + 1)
threshold_freq_lower_idx = int((np.abs(f - config.threshold_freq_lower)).argmin())
threshold_freq_upper_idx = int((np.abs(f - config.threshold_freq_upper)).argmin())
f = f[threshold_freq_lower_idx:threshold_freq_upper_idx]


def install():
    """Upload geometry + band limits (idempotent; called at import and after config changes)."""
    xs = np.ascontiguousarray(x_scan, np.float64)
    ys = np.ascontiguousarray(y_scan, np.float64)
    mx = np.ascontiguousarray(x_i, np.float64)
    my = np.ascontiguousarray(y_i, np.float64)
    act = np.ascontiguousarray(active_mics, np.int32)
    p = _native.ptr
    _native.check(_native.lib().bf_fd_setup(config.N_MICROPHONES, N, float(fs), float(c),
                                            threshold_freq_lower_idx, threshold_freq_upper_idx,
                                            p(xs), len(xs), p(ys), len(ys), float(config.Z),
                                            p(mx), p(my), p(act), len(act)))
