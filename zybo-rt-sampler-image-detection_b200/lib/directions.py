"""`lib.directions` -- delay-table / steering-coefficient generator, B200 edition.

Same public names, arguments and return values as the reference's Cython module
(PC/src/directions.pyx): active_microphones, calc_r_prime, calculate_delays,
calculate_delays_, calculate_delay_miso, get_h, get_h2, compute_convolve_h,
calculate_coefficients.  Sizes come from `interface.config` at call time.

The O(D*n) delay table is evaluated on the GPU (csrc/bf_tables.cu
delay_table_kernel, float64, the reference's operation order, bit-identical);
only the O(X+Y+n) scalars (scan-window axes, microphone coordinates) are
prepared here, reproducing the float32-typed locals Cython gives the reference
(SURVEY.md 7.3-2).  There is no CPU path for the table.

Extras (not in the reference): load_pad_from_geometry / load_lerp_from_geometry
generate the table and install it in the library without a host round trip.
"""
import numpy as np

from interface import config
from . import _native

_unused_mics_file = "unused_mics.npy"


def active_microphones():
    """directions.pyx:35-87 -> (sorted mic ids, count).  Every SKIP_N_MICS-th row and
    column of the 8 x (8*arrays) mosaic, minus `unused_mics.npy` (+64, line 62) when
    that file exists in the cwd."""
    mode = config.SKIP_N_MICS
    n_geo, n_arr = config.GEOMETRY_N_MICS, config.GEOMETRY_N_ARRAYS
    rows = np.arange(0, config.ROWS, mode)
    columns = np.arange(0, config.COLUMNS * n_arr, mode)
    per = config.ROWS * config.COLUMNS
    ids = np.arange(n_geo, dtype=np.float64)
    mosaic = np.hstack([ids[a * per:(a + 1) * per].reshape(config.ROWS, config.COLUMNS)
                        for a in range(n_arr)])
    try:
        unused = np.load(_unused_mics_file)
        unused += 64
    except Exception:  # noqa: BLE001 - the reference swallows everything here too
        unused = []
        print("Will use all microphones")
    picked = [int(mosaic[r, c]) for r in rows for c in columns if mosaic[r, c] not in unused]
    picked = np.sort(picked)
    return picked, len(picked)


def calc_r_prime(d):
    """directions.pyx:17-32 -> float64 [2][n]: x/y of the active microphones; arrays sit
    side by side along -x and the whole row is centred on the origin."""
    n_geo, n_arr = config.GEOMETRY_N_MICS, config.GEOMETRY_N_ARRAYS
    half = d / 2
    pos = np.zeros((2, n_geo))
    k = 0
    for a in range(n_arr):
        a = -a
        for row in range(config.ROWS):
            for col in range(config.COLUMNS):
                pos[0, k] = -col * d - half + a * config.COLUMNS * d + a * 0 + config.COLUMNS * n_arr * half
                pos[1, k] = row * d - config.ROWS * half + half
                k += 1
    pos[0, :] -= n_arr * 0 / 2
    act, _ = active_microphones()
    return pos[:, act]


def _scan_scalars():
    """The scalar prologue of calculate_delays (directions.pyx:91-112) with Cython's
    C-float locals: c, fs, d, alpha, z_scan are float32; fs/c is a float32 division;
    z_scan**2 is powf."""
    c = np.float32(config.PROPAGATION_SPEED)
    fs = np.float32(config.SAMPLE_RATE)
    d = float(np.float32(config.ELEMENT_DISTANCE))
    alpha = float(np.float32(config.VIEW_ANGLE))
    z = np.float32(config.Z)
    aspect = 16 / 9                                   # hard-coded in the reference (line 101)
    k = float(fs / c)
    x_max = float(z) * np.tan((alpha / 2.0) * np.pi / 180)
    y_max = x_max / aspect
    xs = np.linspace(-x_max, x_max, config.MAX_RES_X)
    ys = np.linspace(-y_max, y_max, config.MAX_RES_Y)
    z2 = float(np.float32(z * z))
    return k, xs, ys, z2, d


def _generate(want_f64=False, want_i32=False, want_f32=False, load_algo=-1):
    k, xs, ys, z2, d = _scan_scalars()
    pos = np.ascontiguousarray(calc_r_prime(d))
    n = pos.shape[1]
    X, Y = config.MAX_RES_X, config.MAX_RES_Y
    xs = np.ascontiguousarray(xs, np.float64)
    ys = np.ascontiguousarray(ys, np.float64)
    mx = np.ascontiguousarray(pos[0], np.float64)
    my = np.ascontiguousarray(pos[1], np.float64)
    f64 = np.empty((X, Y, n), np.float64) if want_f64 else None
    i32 = np.empty((X, Y, n), np.int32) if want_i32 else None
    f32 = np.empty((X, Y, n), np.float32) if want_f32 else None
    if load_algo >= 0:
        _native.configure_from(config)
    p = lambda a: _native.ptr(a) if a is not None else None
    _native.check(_native.lib().bf_generate_delays(k, p(xs), X, p(ys), Y, z2, p(mx), p(my), n,
                                                   p(f64), p(i32), p(f32), load_algo))
    return f64, i32, f32


def calculate_delays():
    """directions.pyx:90-124 -> float64 [MAX_RES_X][MAX_RES_Y][n] delays in samples,
    >= 0, the farthest microphone of each direction at 0.  GPU-evaluated."""
    return _generate(want_f64=True)[0]


def calculate_coefficients():
    """directions.pyx:260-277 -> (whole int64 [X][Y][n], taps float32 [X][Y][n][8]).
    `whole` is astype(int) truncation; the 8-tap windowed-sinc table (get_h of the
    fractional part) is what every caller of the reference discards
    (main.pyx:177,209,282,384) -- it is built here in vectorised form, not in a
    12-second Python loop."""
    delays = calculate_delays()
    whole = delays.astype(int)
    frac = delays - whole
    return whole, _get_h_table(frac)


def _get_h_table(delay):
    """get_h (directions.pyx:189-205) over a whole table."""
    tau = -delay[..., None]
    n = np.arange(8)
    x = n - (8 - 1) / 2 - (0.5 + tau) + 1e-9
    h = np.sin(x * np.pi) / (x * np.pi)
    h = h * (0.42 - 0.5 * np.cos(2 * np.pi * n / 8) + 0.08 * np.cos(4 * np.pi * n / 8))
    h = h / np.sum(h, axis=-1, keepdims=True)
    return h.astype(np.float32)


def get_h(delay, N=8):
    """directions.pyx:189-205: 8-tap sinc x Blackman fractional-delay FIR, unity gain."""
    tau = -delay
    n = np.arange(N)
    x = n - (8 - 1) / 2 - (0.5 + tau) + 1e-9
    h = np.sin(x * np.pi) / (x * np.pi)
    h *= 0.42 - 0.5 * np.cos(2 * np.pi * n / 8) + 0.08 * np.cos(4 * np.pi * n / 8)
    h /= np.sum(h)
    return h


def get_h2(delay, N=64):
    """directions.pyx:207-226: N-tap sinc x Blackman FIR for the *whole* delay."""
    return _get_h2_table(np.asarray(delay, dtype=np.float64).reshape(1), N)[0]


def _get_h2_table(delays, T):
    eps = 1e-9
    tau = 0.5 - delays + eps
    taps = np.zeros(delays.shape + (T,), dtype=np.float32)
    total = np.zeros(delays.shape)
    for i in range(T):
        v = i - (T - 1) / 2 - tau
        v = np.sin(v * np.pi) / (v * np.pi)
        w = i * 2 - T + 1
        v = v * (0.42 + 0.5 * np.cos(np.pi * w / (T - 1 + eps))
                 + 0.08 * np.cos(2 * np.pi * w / (T - 1 + eps)))
        total = total + v
        taps[..., i] = v
    return np.divide(taps, total[..., None], dtype=np.float64).astype(np.float32)


def compute_convolve_h():
    """directions.pyx:229-247 -> float32 [X][Y][n][N_TAPS]."""
    delays = calculate_delays()
    print(delays.shape)
    return _get_h2_table(delays, config.N_TAPS)


def calculate_delays_():
    """directions.pyx:126-154: legacy angular generator for one 8x8 array, float32
    [MAX_RES_X][MAX_RES_Y][COLUMNS*ROWS*arrays] (only the first 64 columns are filled)."""
    distance = 0.02
    n_arr = config.GEOMETRY_N_ARRAYS
    R, C = config.ROWS, config.COLUMNS
    out = np.zeros((config.MAX_RES_X, config.MAX_RES_Y, C * R * n_arr), dtype=np.float32)
    half = distance / 2.0
    col = np.arange(C) * distance - C * half + half
    row = np.arange(R) * distance - R * half + half
    for xi, x in enumerate(np.linspace(-config.MAX_ANGLE, config.MAX_ANGLE, config.MAX_RES_X)):
        xf = np.sin(x * -np.pi / 180.0)
        for yi, y in enumerate(np.linspace(-config.MAX_ANGLE, config.MAX_ANGLE, config.MAX_RES_Y)):
            yf = np.sin(y * -np.pi / 180.0)
            t = (col[None, :] * xf + row[:, None] * yf).ravel()
            smallest = min(0, t.min())
            out[xi, yi, :R * C] = t
            out[xi, yi, :] -= smallest
    out *= float(np.float32(config.SAMPLE_RATE) / np.float32(config.PROPAGATION_SPEED))
    return out


def calculate_delay_miso(azimuth, elevation):
    """directions.pyx:156-187: integer delays of one 8x8 array for one angle pair."""
    distance = 0.02
    R, C = config.ROWS, config.COLUMNS
    samp = np.zeros((C * R * config.GEOMETRY_N_ARRAYS), dtype=np.float32)
    xf = np.sin(azimuth * (-np.pi / 180.0))
    yf = np.sin(elevation * (-np.pi / 180.0))
    smallest = 0
    half = distance / 2.0
    for row in range(R):
        for col in range(C):
            t = (col * distance - C * half + half) * xf + (row * distance - R * half + half) * yf
            if t < smallest:
                smallest = t
            samp[row * C + col] = t
    samp -= smallest
    samp *= float(np.float32(config.SAMPLE_RATE) / np.float32(config.PROPAGATION_SPEED))
    return samp.astype(int)


# ---- extras: generate on the device and install without a host round trip ------

def load_pad_from_geometry():
    """== load_coefficients_pad(calculate_coefficients()[0].astype(int32)) of
    main.pyx:176-181, without leaving the GPU."""
    _generate(load_algo=_native.ALGO_PAD)


def load_lerp_from_geometry():
    """== load_coefficients_lerp(float32(calculate_delays())) of main.pyx:332-338."""
    _generate(load_algo=_native.ALGO_LERP)


def whole_and_f32():
    """(int32 whole table, float32 delays) in one generator pass."""
    _, i32, f32 = _generate(want_i32=True, want_f32=True)
    return i32, f32
