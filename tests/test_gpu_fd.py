"""GPU: frequency-domain DAS against the golden heat map produced by importing the
reference's own NumPy module (oracle/gen_golden.py:gen_fd) and against an fp64 NumPy restatement.

Floating-point path: the reference works in complex128 on a complex64 FFT; the kernels work in
fp32 with the phase reduced in fp64.  Tolerance (written here, from north_star): 1e-5 relative per
pixel, and 1e-5 of the map's maximum (the heat map is normalised by its own maximum)."""
import numpy as np
import pytest

from util import gold

pytestmark = pytest.mark.gpu

TOL_MAX_REL = 1e-5       # |a-b| <= 1e-5 * max|b|
TOL_PIXEL_REL = 1e-5


def _fd_numpy(sig_nm, g):
    """fp64 restatement of beam_forming_algorithm.py:30-63 from the fixture's geometry."""
    X = np.fft.rfft(sig_nm.astype(np.float64), axis=0)[int(g["lo"]):int(g["hi"])]      # (F, M)
    xs, ys, f = g["x_scan"], g["y_scan"], g["f"]
    mx, my = g["r_prime_all"]
    r = np.sqrt(xs[:, None] ** 2 + ys[None, :] ** 2 + 1.0)
    u = (xs[:, None, None] * mx + ys[None, :, None] * my) / r[:, :, None]              # (X, Y, M)
    k = 2 * np.pi * f / float(g["c"])
    P = np.zeros(u.shape[:2])
    for i in range(len(f)):
        s = (X[i][None, None, :] * np.exp(-1j * k[i] * u)).sum(-1)
        P += np.abs(s) ** 2
    return P


def test_fd_das_matches_reference_module():
    g = gold("fd_das")
    from realtime_scripts import beam_forming_algorithm as bfa, calc_phase_shift_cartesian as cps
    assert (cps.threshold_freq_lower_idx, cps.threshold_freq_upper_idx) == (int(g["lo"]), int(g["hi"]))
    assert np.array_equal(cps.x_scan, g["x_scan"]) and np.array_equal(cps.y_scan, g["y_scan"])
    assert np.array_equal(np.stack([cps.x_i, cps.y_i]), g["r_prime_all"]) and np.array_equal(cps.f, g["f"])
    heat = bfa.main(g["signal"])
    ref = g["heatmap"]
    assert heat.shape == ref.shape == (13, 13) and heat.dtype == np.float64
    assert np.unravel_index(heat.argmax(), heat.shape) == np.unravel_index(ref.argmax(), ref.shape) == (9, 4)
    err = np.abs(heat - ref)
    assert err.max() <= TOL_MAX_REL * ref.max(), err.max()
    assert np.all(err <= TOL_PIXEL_REL * np.abs(ref) + 1e-12), (err / np.abs(ref)).max()
    # un-normalised power against the reference's own intermediate and the fp64 restatement
    P = bfa.power(g["signal"]).astype(np.float64)
    assert np.abs(P - g["fft_power"]).max() <= TOL_MAX_REL * g["fft_power"].max()
    P64 = _fd_numpy(g["signal"], g)
    assert np.abs(P - P64).max() <= TOL_MAX_REL * P64.max()
    # a 1e-4 times weaker input is still above the 0.2 threshold here: same normalised map
    weak = bfa.main((g["signal"] * 1e-4).astype(np.float32))
    assert np.abs(weak - g["heatmap_quiet"]).max() <= TOL_MAX_REL * g["heatmap_quiet"].max()
    # below the threshold (max of the un-normalised power < 0.2) the reference returns all zeros
    scale = np.float32(0.1 * np.sqrt(0.2 / P64.max()))
    assert (P64.max() * float(scale) ** 2) < 0.2
    quiet = bfa.main((g["signal"] * scale).astype(np.float32))
    assert quiet.shape == (13, 13) and not quiet.any()


def test_fd_linearity_and_batch():
    """Size-independent properties: power scales with the square of the input; batched device
    call equals per-frame calls."""
    import torch
    g = gold("fd_das")
    from lib import _native as nat
    from realtime_scripts import beam_forming_algorithm as bfa
    sig = g["signal"]
    P1 = bfa.power(sig)
    P2 = bfa.power((sig * np.float32(2)).astype(np.float32))
    assert np.array_equal(P2, P1 * np.float32(4))
    rng = np.random.default_rng(4)
    frames = rng.standard_normal((3, 256, 256)).astype(np.float32)
    frames[1] = sig.T
    L = nat.lib()
    d_sig = torch.from_numpy(frames).cuda()
    d_heat = torch.zeros((3, 169), device="cuda")
    nat.check(L.bf_fd_das_dev(d_sig.data_ptr(), d_heat.data_ptr(), 3, 0.2, 0, None))
    torch.cuda.synchronize()
    got = d_heat.cpu().numpy()
    assert np.array_equal(got[1].reshape(13, 13), P1)
    for i in (0, 2):
        assert np.array_equal(got[i].reshape(13, 13), bfa.power(frames[i].T))
