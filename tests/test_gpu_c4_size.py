"""GPU: the frequency-domain path at BASELINE config C4's size -- 256 microphones, 1024-point FFT,
bins 1..512, K = 64 snapshots, 256 x 128 = 32 768 directions -- against the float64 NumPy oracle.

PARITY UNPINNED for MVDR (the reference has none); the oracle restates the definition with the
reference's conventions (real FFT along samples, no window, no scaling:
beam_forming_algorithm.py:31-33; steering phase exp(-j k_f (x_s x_m + y_s y_m)/r_s):
calc_phase_shift_cartesian.py:44-48; frequency axis linspace(0, int(fs/2), N/2+1)).  The oracle is
evaluated on a SAMPLE of the map -- both edge directions of every 128-direction MMA tile plus
512 random ones -- and on the full covariance of 8 bins; the CUDA path computes everything.

Tolerances (floating point, stated): covariance 1e-12 of its maximum (float64 end to end);
MVDR and DAS power 1e-5 relative PER PIXEL (north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, N, F, K, RES_X, RES_Y = 256, 1024, 512, 64, 256, 128
D = RES_X * RES_Y
FS, C, LO = 48828.0, 343.0, 1
LOADING = 1e-2
REL_TOL = 1e-5


def _geometry():
    import realtime_scripts.calc_r_prime as rp
    import realtime_scripts.config as cfg
    pos, _ = rp.calc_r_prime(cfg.ELEMENT_DISTANCE)
    x_max = np.tan(np.deg2rad(cfg.VIEW_ANGLE / 2))
    xs = np.linspace(-x_max, x_max, RES_X)
    ys = np.linspace(-x_max / cfg.ASPECT_RATIO, x_max / cfg.ASPECT_RATIO, RES_Y)
    return xs, ys, np.ascontiguousarray(pos[0]), np.ascontiguousarray(pos[1])


def _snapshots(xs, ys, mx, my, seed=1237):
    """K non-overlapping 1024-sample frames of three far-field sources (grid cells) + noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(N)
    snaps = rng.normal(0.0, 0.05, (K, M, N))
    for (ix, iy, f0, amp) in ((200, 40, 2000.0, 0.3), (60, 100, 5000.0, 0.2), (128, 64, 9000.0, 0.1)):
        r = np.sqrt(xs[ix] ** 2 + ys[iy] ** 2 + 1.0)
        tau = (xs[ix] * mx + ys[iy] * my) / r / C * FS                      # samples
        ph = rng.uniform(0, 2 * np.pi, (K, 1, 1))
        snaps += amp * np.sin(2 * np.pi * f0 * (t[None, None, :] + tau[None, :, None]) / FS + ph)
    return snaps.astype(np.float32)


def _sample_dirs(seed=3):
    edges = np.concatenate([np.arange(0, D, 128), np.arange(127, D, 128)])
    rnd = np.random.default_rng(seed).choice(D, 512, replace=False)
    return np.unique(np.concatenate([edges, rnd]))


def _steering_u(xs, ys, mx, my, dirs):
    ix, iy = dirs // RES_Y, dirs % RES_Y                       # flat index d = x * RES_Y + y
    r = np.sqrt(xs[ix] ** 2 + ys[iy] ** 2 + 1.0)
    return (xs[ix, None] * mx[None, :] + ys[iy, None] * my[None, :]) / r[:, None]     # (S, M)


def _freqs():
    return np.linspace(0, int(FS / 2), N // 2 + 1)[LO:LO + F]


@pytest.fixture(scope="module")
def c4():
    from lib import _native as nat
    L = nat.lib()
    xs, ys, mx, my = _geometry()
    act = np.arange(M, dtype=np.int32)
    p = nat.ptr
    nat.check(L.bf_fd_setup(M, N, FS, C, LO, LO + F, p(xs), RES_X, p(ys), RES_Y, 1.0, p(mx), p(my), p(act), M))
    snaps = _snapshots(xs, ys, mx, my)
    dirs = _sample_dirs()
    u = _steering_u(xs, ys, mx, my, dirs)
    X = np.fft.rfft(snaps.astype(np.float64), axis=2)[:, :, LO:LO + F]          # (K, M, F)
    X = np.ascontiguousarray(np.transpose(X, (2, 0, 1)))                        # (F, K, M)
    yield dict(nat=nat, L=L, snaps=snaps, dirs=dirs, u=u, X=X, k=2 * np.pi * _freqs() / C)
    from realtime_scripts import beam_forming_algorithm as bfa               # the stock FD geometry is
    bfa._installed = False                                                  # re-installed on next use


def test_mvdr_c4_size_vs_float64_oracle(c4):
    nat, L, X, u, k, dirs = c4["nat"], c4["L"], c4["X"], c4["u"], c4["k"], c4["dirs"]
    P = np.zeros(D, np.float32)
    nat.check(L.bf_fd_mvdr(nat.ptr(c4["snaps"]), nat.ptr(P), K, LOADING))
    assert np.isfinite(P).all() and (P > 0).all()
    # ---- oracle: covariance + loading, inverse, quadratic form, float64 -------------------------
    R = np.matmul(np.transpose(X, (0, 2, 1)), X.conj()) / K                     # (F, M, M): sum_k x x^H
    tr = np.einsum("fii->f", R).real
    R += (LOADING * tr / M)[:, None, None] * np.eye(M)[None]
    P_ref = np.zeros(len(dirs))
    for f in range(F):
        A = np.exp(-1j * k[f] * u)                                              # (S, M), rows a^T
        Rinv = np.linalg.inv(R[f])
        P_ref += 1.0 / np.einsum("sm,sm->s", A.conj() @ Rinv, A).real
    err = np.abs(P[dirs] - P_ref) / P_ref
    print("mvdr C4 size: %d sampled directions, max per-pixel rel err %.3e (mean %.3e)" % (len(dirs), err.max(), err.mean()))
    assert err.max() <= REL_TOL
    # ---- full covariance of 8 bins (first, last, and six in between) ----------------------------
    cov = np.zeros((F, M, M, 2))
    nat.check(L.bf_fd_get_covariance(nat.ptr(cov), F * M * M))
    for f in (0, 1, 63, 64, 200, 333, 510, 511):
        Rg = cov[f, ..., 0] + 1j * cov[f, ..., 1]
        assert np.abs(Rg - R[f]).max() <= 1e-12 * np.abs(R[f]).max(), f
    # same brightest sampled direction
    assert P[dirs].argmax() == P_ref.argmax()


def test_fd_das_c4_size_vs_float64_oracle(c4):
    import torch
    nat, L, X, u, k, dirs = c4["nat"], c4["L"], c4["X"], c4["u"], c4["k"], c4["dirs"]
    frames = c4["snaps"][:2]                                                     # two frames, (M, N) each
    d_sig = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((2, D), device="cuda")
    nat.check(L.bf_fd_das_dev(d_sig.data_ptr(), d_out.data_ptr(), 2, 0.0, 0, None))     # un-normalised power
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().astype(np.float64)
    for i in range(2):
        P_ref = np.zeros(len(dirs))
        for f in range(F):
            s = np.exp(-1j * k[f] * u) @ X[f, i]                                # sum_m X[f, m] e^{-j k u}
            P_ref += np.abs(s) ** 2
        err = np.abs(got[i, dirs] - P_ref) / P_ref
        print("fd-das C4 size frame %d: max per-pixel rel err %.3e" % (i, err.max()))
        assert err.max() <= REL_TOL
