"""NumPy restatement of the reference delay-table generator (PC/src/directions.pyx).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  Parity status: PINNED --
tests/test_oracle_vs_ref.py compares every function bitwise with the
cythonized reference module in oracle/_ref/<cfg>/lib/directions*.so, and
tests/test_golden.py with the committed tests/golden/ fixtures.

The restatement is parametrised by a plain dict of the config.json "general"
constants instead of compile-time macros.  Two literals of the reference ignore
config.json (directions.pyx:15-16: _N_MICS = 256, _ACTIVE_MICS = 4); they are the
`n_mics_geom` / `n_arrays_geom` keys here (64 / 1 for BASELINE config C1).

Cython types the locals c, fs, d, alpha, z_scan of calculate_delays() as C
`float` (they are assigned from `float` externs in config.pxd:14-23), so
  * d      = (double)(float)0.02      (passed on to calc_r_prime as a Python float)
  * fs/c   is a float32 division, then promoted
  * alpha/2 is float / double -> double
  * z_scan**2 is powf(z, 2.0f)
Everything else is float64 NumPy in the source's operation order.
All citations: /root/reference/PC/src/directions.pyx.
"""
import numpy as np

DEFAULTS = dict(
    N_MICROPHONES=256, N_SAMPLES=256, N_TAPS=8, COLUMNS=8, ROWS=8,
    MAX_RES_X=57, MAX_RES_Y=32, Z=1.0, MAX_ANGLE=70.0, VIEW_ANGLE=59.0,
    SAMPLE_RATE=48828.0, ELEMENT_DISTANCE=0.02, SKIP_N_MICS=1,
    PROPAGATION_SPEED=340.0, n_mics_geom=256, n_arrays_geom=4,
)


def cfg_with(**kw):
    c = dict(DEFAULTS)
    c.update(kw)
    return c


def active_microphones(cfg, unused_mics=None):
    """directions.pyx:35-87.  `unused_mics` replaces the np.load('unused_mics.npy')
    of line 61 (the reference adds 64 to the loaded ids, line 62)."""
    rows_n, cols_n = cfg["ROWS"], cfg["COLUMNS"]
    nm, na, mode = cfg["n_mics_geom"], cfg["n_arrays_geom"], cfg["SKIP_N_MICS"]
    rows = np.arange(0, rows_n, mode)
    columns = np.arange(0, cols_n * na, mode)
    mics = np.linspace(0, nm - 1, nm)
    per = rows_n * cols_n
    mosaic = np.linspace(0, per - 1, per).reshape((rows_n, cols_n))
    for a in range(1, na):
        mosaic = np.hstack((mosaic, mics[a * per:(a + 1) * per].reshape((rows_n, cols_n))))
    banned = [] if unused_mics is None else list(np.asarray(unused_mics) + 64)
    chosen = []
    for r in rows:
        for c in columns:
            mic = mosaic[r, c]
            if mic not in banned:
                chosen.append(int(mic))
    chosen = np.sort(chosen)
    return chosen, len(chosen)


def calc_r_prime(cfg, d, unused_mics=None):
    """directions.pyx:17-32: arrays side by side along -x, centred."""
    rows_n, cols_n = cfg["ROWS"], cfg["COLUMNS"]
    nm, na = cfg["n_mics_geom"], cfg["n_arrays_geom"]
    half = d / 2
    pos = np.zeros((2, nm))
    k = 0
    for array in range(na):
        array *= -1
        for row in range(rows_n):
            for col in range(cols_n):
                pos[0, k] = -col * d - half + array * cols_n * d + array * 0 + cols_n * na * half
                pos[1, k] = row * d - rows_n * half + half
                k += 1
    pos[0, :] -= na * 0 / 2
    act, _ = active_microphones(cfg, unused_mics)
    return pos[:, act]


def scan_axes(cfg):
    """directions.pyx:91-112: the scan-window axes and the scale factor, with
    the float32-typed locals reproduced.  Returns (k, x_scan[X], y_scan[Y], z2)."""
    c = np.float32(cfg["PROPAGATION_SPEED"])
    fs = np.float32(cfg["SAMPLE_RATE"])
    alpha = np.float32(cfg["VIEW_ANGLE"])
    z = np.float32(cfg["Z"])
    AS = 16 / 9
    k = float(fs / c)                       # float32 division, then promoted
    zf = float(z)
    x_max = zf * np.tan((float(alpha) / 2.0) * np.pi / 180)
    y_max = x_max / AS
    xs = np.linspace(-x_max, x_max, cfg["MAX_RES_X"])
    ys = np.linspace(-y_max, y_max, cfg["MAX_RES_Y"])
    z2 = float(np.float32(z * z))           # powf(z_scan, 2.0f)
    return k, xs, ys, z2


def calculate_delays(cfg, unused_mics=None):
    """directions.pyx:90-124 -> float64 [MAX_RES_X][MAX_RES_Y][n]."""
    d = float(np.float32(cfg["ELEMENT_DISTANCE"]))
    pos = calc_r_prime(cfg, d, unused_mics)
    xi, yi = pos[0, :], pos[1, :]
    k, xs, ys, z2 = scan_axes(cfg)
    X, Y = cfg["MAX_RES_X"], cfg["MAX_RES_Y"]
    xs = xs.reshape(X, 1, 1)
    ys = ys.reshape(1, Y, 1)
    r = np.sqrt(xs ** 2 + ys ** 2 + z2)
    delay = k * (xs * xi + ys * yi) / r
    delay -= np.amin(delay, axis=2).reshape(X, Y, 1)
    return delay


def whole_samples(cfg, unused_mics=None):
    """directions.pyx:262-265: astype(int) truncation (first return value of
    calculate_coefficients; the FIR taps it also builds are discarded by every
    caller, main.pyx:177,209,282,384)."""
    return calculate_delays(cfg, unused_mics).astype(int)


def get_h(delay, N=8):
    """directions.pyx:189-205 (8-tap windowed sinc; literal 8 in window/centre)."""
    tau = -delay
    epsilon = 1e-9
    n = np.arange(N)
    x = n - (8 - 1) / 2 - (0.5 + tau) + epsilon
    h = np.sin(x * np.pi) / (x * np.pi)
    win = 0.42 - 0.5 * np.cos(2 * np.pi * n / 8) + 0.08 * np.cos(4 * np.pi * n / 8)
    h *= win
    h /= np.sum(h)
    return h


def get_h2(delay, N=64):
    """directions.pyx:207-226 (N-tap sinc x Blackman, float32 storage of each
    tap, float64 running sum of the un-rounded taps)."""
    epsilon = 1e-9
    tau = 0.5 - delay + epsilon
    h = np.zeros(N, dtype=np.float32)
    total = 0
    for i in range(N):
        hi = i - (N - 1) / 2 - tau
        hi = np.sin(hi * np.pi) / (hi * np.pi)
        n = i * 2 - N + 1
        black = (0.42 + 0.5 * np.cos(np.pi * n / (N - 1 + epsilon))
                 + 0.08 * np.cos(2 * np.pi * n / (N - 1 + epsilon)))
        hi *= black
        total += hi
        h[i] = hi
    h /= total
    return h


def compute_convolve_h(cfg, unused_mics=None):
    """directions.pyx:229-247 -> float32 [X][Y][n][N_TAPS] (vectorised over the
    table; same per-tap arithmetic as get_h2)."""
    delays = calculate_delays(cfg, unused_mics)
    T = cfg["N_TAPS"]
    epsilon = 1e-9
    tau = 0.5 - delays + epsilon
    out = np.zeros(delays.shape + (T,), dtype=np.float32)
    total = np.zeros(delays.shape)
    for i in range(T):
        hi = i - (T - 1) / 2 - tau
        hi = np.sin(hi * np.pi) / (hi * np.pi)
        n = i * 2 - T + 1
        black = (0.42 + 0.5 * np.cos(np.pi * n / (T - 1 + epsilon))
                 + 0.08 * np.cos(2 * np.pi * n / (T - 1 + epsilon)))
        hi = hi * black
        total = total + hi
        out[..., i] = hi
    # h /= sum_ : float32 array divided in place by a float64 scalar
    out = np.divide(out, total[..., None], dtype=np.float64).astype(np.float32)
    return out


def calculate_delay_miso(cfg, azimuth, elevation):
    """directions.pyx:156-187 legacy angular single-array delays (int)."""
    distance = 0.02
    rows_n, cols_n, na = cfg["ROWS"], cfg["COLUMNS"], cfg["n_arrays_geom"]
    samp = np.zeros((cols_n * rows_n * na), dtype=np.float32)
    azimuth = azimuth * (-np.pi / 180.0)
    xf = np.sin(azimuth)
    elevation = elevation * (-np.pi / 180.0)
    yf = np.sin(elevation)
    smallest = 0
    for row in range(rows_n):
        for col in range(cols_n):
            half = distance / 2.0
            tc = col * distance - cols_n * half + half
            tr = row * distance - rows_n * half + half
            t = tc * xf + tr * yf
            if t < smallest:
                smallest = t
            samp[row * cols_n + col] = t
    samp -= smallest
    # SAMPLE_RATE / PROPAGATION_SPEED is a C float division (config.pxd:15,21)
    samp *= float(np.float32(cfg["SAMPLE_RATE"]) / np.float32(cfg["PROPAGATION_SPEED"]))
    return samp.astype(int)


def steer_offset_degree(cfg, azimuth, elevation, n_active):
    """main.pyx:498-513 steer_cartesian_degree arithmetic -> table row offset."""
    assert -90 <= azimuth <= 90 and -90 <= elevation <= 90
    az = int(((azimuth + 90) / 180) * cfg["MAX_RES_X"])
    el = int(((elevation + 90) / 180) * cfg["MAX_RES_Y"])
    return int(el * cfg["MAX_RES_X"] * n_active + az * n_active)


def steer_offset_unit(cfg, x, y, n_active):
    """main.pyx:515-528 stear_miso_beam arithmetic."""
    az = int(x * cfg["MAX_RES_X"])
    el = int(y * cfg["MAX_RES_Y"])
    return int(el * cfg["MAX_RES_X"] * n_active + az * n_active)
