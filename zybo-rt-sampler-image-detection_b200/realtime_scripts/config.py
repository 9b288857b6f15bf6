"""Constants of the frequency-domain backend (PC/application/realtime_scripts/config.py: a stale,
hand-edited copy of the generated config with its own values -- 13x13 grid, c = 343 m/s, 16/9,
ACTIVE_ARRAYS = 4, band 0-18 kHz).  Plain assignments; edit or override before importing
calc_phase_shift_cartesian, exactly like the reference."""
import ctypes

import numpy

N_MICROPHONES = 256
N_SAMPLES = 256
COLUMNS = 8
ROWS = 8
MAX_RES_X = 13
MAX_RES_Y = 13
Z = 1.0
VIEW_ANGLE = 68.0
ELEMENT_DISTANCE = 0.02
ARRAY_SEPARATION = 0.0
ACTIVE_ARRAYS = 4
PROPAGATION_SPEED = 343.0
ASPECT_RATIO = 16 / 9
SAMPLE_RATE = 48828.0
columns = 8
rows = 8
mode = 1
plot_setup = 0
threshold_freq_lower = 0
threshold_freq_upper = 18000
fs = int(48828)
DTYPE = ctypes.c_int32
NP_DTYPE = numpy.float32
