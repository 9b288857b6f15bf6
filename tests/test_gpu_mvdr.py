"""GPU: frequency-domain MVDR.  PARITY UNPINNED -- the reference has no MVDR; the
oracle is the float64 NumPy restatement below, which takes its FFT/bin convention from
beam_forming_algorithm.py:31-33 and its steering phase from calc_phase_shift_cartesian.py:44-48.
The reference-pinned anchor: the DAS form a^H R a of the SAME covariance must reproduce the
frequency-domain DAS power (a19 path, itself checked against the reference module's golden map).

Tolerance (stated, floating point): covariance 1e-12 relative (float64 end to end); MVDR map
1e-5 relative per pixel (measured on B200: 9e-7 with the CUDA-core fp32 contraction), against a
float64 oracle at cond(R) up to M/loading = 2.6e4."""
import numpy as np
import pytest

from util import gold

pytestmark = pytest.mark.gpu


def _setup():
    from realtime_scripts import beam_forming_algorithm as bfa
    bfa._ensure()
    from lib import _native as nat
    return bfa, nat, nat.lib()


def _snapshots(g, K, seed):
    """K snapshots: three sources at grid cells (9,4), (3,10), (6,6) + noise, (K, M, N) float32."""
    rng = np.random.default_rng(seed)
    xs, ys = g["x_scan"], g["y_scan"]
    mx, my = g["r_prime_all"]
    N, M = 256, 256
    t = np.arange(N)
    snaps = np.zeros((K, M, N))
    for (ix, iy, f0, amp) in ((9, 4, 3000.0, 0.2), (3, 10, 5200.0, 0.15), (6, 6, 7100.0, 0.1)):
        r = np.sqrt(xs[ix] ** 2 + ys[iy] ** 2 + 1.0)
        tau = (xs[ix] * mx + ys[iy] * my) / r / float(g["c"]) * float(g["fs"])
        for k in range(K):
            ph = rng.uniform(0, 2 * np.pi)
            snaps[k] += amp * np.sin(2 * np.pi * f0 * (t[None, :] + tau[:, None]) / float(g["fs"]) + ph)
    snaps += rng.normal(0, 0.02, snaps.shape)
    return snaps.astype(np.float32)


def _oracle(snaps, g, loading):
    lo, hi = int(g["lo"]), int(g["hi"])
    X = np.fft.rfft(snaps.astype(np.float64), axis=2)[:, :, lo:hi]            # (K, M, F)
    X = np.transpose(X, (2, 0, 1))                                             # (F, K, M)
    K, M = X.shape[1], X.shape[2]
    R = np.einsum("fki,fkj->fij", X, X.conj()) / K
    tr = np.einsum("fii->f", R).real
    R = R + (loading * tr / M)[:, None, None] * np.eye(M)[None]
    xs, ys, f = g["x_scan"], g["y_scan"], g["f"]
    mx, my = g["r_prime_all"]
    r = np.sqrt(xs[:, None] ** 2 + ys[None, :] ** 2 + 1.0)
    u = ((xs[:, None, None] * mx + ys[None, :, None] * my) / r[:, :, None]).reshape(-1, M)   # (D, M)
    k = 2 * np.pi * f / float(g["c"])
    P = np.zeros(u.shape[0])
    Pdas = np.zeros(u.shape[0])
    for i in range(len(f)):
        A = np.exp(-1j * k[i] * u)                                             # (D, M) rows a^T
        Rinv = np.linalg.inv(R[i])
        q = np.einsum("dm,mn,dn->d", A.conj(), Rinv, A).real
        P += 1.0 / q
        Rd = R[i] - (loading * tr[i] / M) * np.eye(M)
        # DAS through the covariance: mean_k |sum_m X_k[m] a[m]|^2 = a^T R conj(a) (a19 convention)
        Pdas += np.einsum("dm,mn,dn->d", A, Rd, A.conj()).real
    return R, P, Pdas


@pytest.mark.parametrize("tensor_cores", [0, 1, 2, 3, 4])
def test_mvdr_against_float64_oracle_and_das_anchor(tensor_cores, monkeypatch):
    """tensor_cores=4 (default): as 3 with N = 128 MMAs and quarter-granular triangular skipping;
    tensor_cores=3: warp-specialised tcgen05 steering contraction, kind::f16 with a two-term
    fp16 split; 2: the same with kind::tf32 (3-pass split tf32); 1: its single-buffered first version;
    0: CUDA-core fp32."""
    monkeypatch.setenv("BF_MVDR_TC", str(tensor_cores))
    g = gold("fd_das")
    bfa, nat, L = _setup()
    K, loading = 12, 1e-2
    snaps = _snapshots(g, K, 5)
    P = np.zeros(169, np.float32)
    nat.check(L.bf_fd_mvdr(nat.ptr(snaps), nat.ptr(P), K, loading))
    R_ref, P_ref, Pdas_ref = _oracle(snaps, g, loading)
    # (1) covariance with diagonal loading, float64 end to end
    F, M = R_ref.shape[0], 256
    cov = np.zeros((F, M, M, 2))
    nat.check(L.bf_fd_get_covariance(nat.ptr(cov), F * M * M))
    R = cov[..., 0] + 1j * cov[..., 1]
    assert np.abs(R - R_ref).max() <= 1e-12 * np.abs(R_ref).max()
    # (2) reference-pinned anchor: DAS power through R == mean over snapshots of the a19 path
    das = np.mean([bfa.power(snaps[k].T).astype(np.float64).ravel() for k in range(K)], axis=0)
    assert np.abs(das - Pdas_ref).max() <= 1e-5 * Pdas_ref.max()
    # (3) the MVDR map itself
    err = np.abs(P - P_ref)
    print("mvdr: max err / max = %.2e, max pixel rel = %.2e" % (err.max() / P_ref.max(), (err / P_ref).max()))
    assert err.max() <= 1e-5 * P_ref.max()
    assert np.all(err <= 1e-5 * P_ref)
    # same peak cell as the oracle, and it is one of the three source cells
    assert P.argmax() == P_ref.argmax()
    assert np.unravel_index(P.argmax(), (13, 13)) in ((9, 4), (3, 10), (6, 6))
    # (4) properties: scaling the data by 2 scales P by 4 (loading is relative to the trace)
    P2 = np.zeros(169, np.float32)
    nat.check(L.bf_fd_mvdr(nat.ptr((snaps * np.float32(2)).astype(np.float32)), nat.ptr(P2), K, loading))
    assert np.allclose(P2, 4 * P, rtol=2e-5)


def test_mvdr_rejects_singular_covariance():
    g = gold("fd_das")
    bfa, nat, L = _setup()
    snaps = _snapshots(g, 2, 6)
    P = np.zeros(169, np.float32)
    assert L.bf_fd_mvdr(nat.ptr(snaps), nat.ptr(P), 2, 0.0) != 0        # K < M and no loading
    assert b"positive definite" in L.bf_last_error()


def test_direction_slices_equal_the_full_map():
    """Direction sharding of the frequency-domain maps (FD path) on one GPU: steering only a slice
    of the grid (bf_fd_mvdr_dev_slice / bf_fd_das_dev_slice) gives exactly the values of the full map, whatever
    the slice boundaries (they cut through the 128-direction MMA tile)."""
    import torch
    g = gold("fd_das")
    bfa, nat, L = _setup()
    K, D = 6, 169
    snaps = torch.from_numpy(_snapshots(g, K, 9)).cuda()
    full = torch.zeros(D, device="cuda")
    nat.check(L.bf_fd_mvdr_dev(snaps.data_ptr(), full.data_ptr(), K, 1e-2, None))
    das = torch.zeros((2, D), device="cuda")
    nat.check(L.bf_fd_das_dev(snaps.data_ptr(), das.data_ptr(), 2, 0.0, 0, None))
    torch.cuda.synchronize()
    for bounds in ((0, 100, 169), (0, 1, 128, 129, 169)):
        parts, dparts = [], []
        for a, b in zip(bounds[:-1], bounds[1:]):
            part = torch.full((b - a,), float("nan"), device="cuda")
            nat.check(L.bf_fd_mvdr_dev_slice(snaps.data_ptr(), part.data_ptr(), K, 1e-2, a, b - a, None))
            parts.append(part)
            dp = torch.full((2, b - a), float("nan"), device="cuda")
            nat.check(L.bf_fd_das_dev_slice(snaps.data_ptr(), dp.data_ptr(), 2, a, b - a, None))
            dparts.append(dp)
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(parts), full), bounds
        assert torch.equal(torch.cat(dparts, dim=1), das), bounds
    assert L.bf_fd_mvdr_dev_slice(snaps.data_ptr(), full.data_ptr(), K, 1e-2, 100, 100, None) != 0     # beyond the grid
