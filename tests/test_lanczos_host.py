"""Host-side control of the eigen-solver of N4 (tadbit_b200/csrc/lanczos.hpp: Lanczos with full
re-orthogonalisation, break-down / restart, selection by magnitude, descending order) driven over
plain arrays (tests/native/lanczos_host.cpp, compiled here with g++) and the exported tridiagonal
QL solver (tb_tridiag_eig, host only), against numpy.linalg.eigh.  The product runs the same
control code with CUDA kernels as its vector operations (tests/test_gpu_compartments.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib():
    src = os.path.join(HERE, "native", "lanczos_host.cpp")
    out_dir = os.path.join(HERE, "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblanczos_host.so")
    hdr = os.path.join(HERE, "..", "tadbit_b200", "csrc", "lanczos.hpp")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    lib = C.CDLL(so)
    lib.lanczos_host.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    lib.lanczos_host.restype = C.c_int
    return lib


def run(lib, a, k, m_cap=0, tol=1e-13):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    kk = min(k, n)
    evals = np.zeros(kk)
    evecs = np.zeros((kk, n))
    info = np.zeros(5, dtype=np.int32)
    lib.lanczos_host(n, a.ctypes.data, k, m_cap, tol, evals.ctypes.data, evecs.ctypes.data,
                     info.ctypes.data)
    return evals[:info[0]], evecs[:info[0]], info


def reference_largest(a, k):
    """What the reference takes from eigsh(k): the k eigenpairs of largest magnitude, read in
    descending algebraic order (k >= n: everything, from eigh)."""
    w, v = np.linalg.eigh(a)
    n = a.shape[0]
    if k >= n:
        idx = np.argsort(-w)
    else:
        idx = np.argsort(-np.abs(w), kind="stable")[:k]
        idx = idx[np.argsort(-w[idx])]
    return w[idx], v[:, idx].T


def check_pairs(a, evals, evecs, k, vec_tol=1e-8):
    w, v = reference_largest(a, k)
    assert evals.shape == w.shape
    scale = max(np.abs(w).max(), 1e-300)
    np.testing.assert_allclose(evals, w, rtol=0, atol=1e-11 * scale)
    for i in range(len(w)):
        # residual of the pair itself and agreement up to sign where the eigenvalue is simple
        r = a @ evecs[i] - evals[i] * evecs[i]
        assert np.linalg.norm(r) <= 1e-10 * scale
        assert abs(np.linalg.norm(evecs[i]) - 1) < 1e-12
        gaps = np.abs(np.linalg.eigvalsh(a) - w[i])
        gaps = np.sort(gaps)
        if len(gaps) > 1 and gaps[1] > 1e-3 * scale:
            assert min(np.abs(evecs[i] - v[i]).max(), np.abs(evecs[i] + v[i]).max()) < vec_tol


def corr_like(n, seed):
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(n, 6)) @ rng.normal(size=(6, n)) + 0.3 * rng.normal(size=(n, n))
    return np.corrcoef(base + base.T)


@pytest.mark.parametrize("n,k,seed", [(60, 3, 1), (200, 3, 2), (500, 5, 3), (37, 1, 4), (150, 10, 5)])
def test_correlation_matrices(host_lib, n, k, seed):
    a = corr_like(n, seed)
    evals, evecs, info = run(host_lib, a, k)
    assert info[3] == 1 and info[0] == k
    assert info[1] < n or n <= 60
    check_pairs(a, evals, evecs, k)


def test_indefinite_matrix_selects_by_magnitude(host_lib):
    rng = np.random.default_rng(7)
    q, _ = np.linalg.qr(rng.normal(size=(80, 80)))
    w = np.concatenate([[-9.0, 7.5, -6.0, 5.0], rng.uniform(-1, 1, 76)])
    a = (q * w) @ q.T
    a = (a + a.T) / 2
    evals, evecs, info = run(host_lib, a, 3)
    assert info[3] == 1
    np.testing.assert_allclose(evals, [7.5, -6.0, -9.0], atol=1e-11)     # |.| largest, descending value
    check_pairs(a, evals, evecs, 3)


@pytest.mark.parametrize("n", [3, 4, 5])
def test_all_pairs_when_k_reaches_n(host_lib, n):
    a = corr_like(n, 11 + n)
    evals, evecs, info = run(host_lib, a, 3 if n == 3 else n)
    assert info[0] == n and info[3] == 1
    check_pairs(a, evals, evecs, n)


def test_low_rank_and_degenerate(host_lib):
    rng = np.random.default_rng(5)
    u = rng.normal(size=40)
    a = np.outer(u, u)                                   # rank one: break-down after the first step
    evals, evecs, info = run(host_lib, a, 3)
    assert info[3] == 1 and info[2] >= 1
    np.testing.assert_allclose(evals[0], u @ u, rtol=1e-12)
    np.testing.assert_allclose(evals[1:], 0.0, atol=1e-10)
    ident = np.eye(12)
    evals, evecs, info = run(host_lib, ident, 3)
    assert info[3] == 1
    np.testing.assert_allclose(evals, 1.0, atol=1e-12)
    np.testing.assert_allclose(evecs @ evecs.T, np.eye(3), atol=1e-10)
    zero = np.zeros((9, 9))
    evals, evecs, info = run(host_lib, zero, 3)
    assert info[3] == 1
    np.testing.assert_allclose(evals, 0.0, atol=1e-300)


def test_block_diagonal_with_zero_rows(host_lib):
    # a correlation matrix whose constant rows became NaN -> 0 (hic_data.py:881)
    a = corr_like(90, 21)
    a[[3, 50], :] = 0.0
    a[:, [3, 50]] = 0.0
    evals, evecs, info = run(host_lib, a, 3)
    assert info[3] == 1
    check_pairs(a, evals, evecs, 3)


def test_step_cap_reports_no_convergence(host_lib):
    a = corr_like(300, 33)
    evals, evecs, info = run(host_lib, a, 3, m_cap=4)
    assert info[3] == 0 and info[1] == 4


def test_tridiagonal_solver_exported_by_the_library():
    from tadbit_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 17, 120):
        d = rng.normal(size=n)
        e = rng.normal(size=max(n - 1, 1))
        if n > 4:
            e[n // 2] = 0.0                              # a split matrix
        t = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
        dd = d.copy()
        vecs = np.zeros((n, n))
        _capi.check(lib.tb_tridiag_eig(n, _capi.ptr(dd, C.c_double), _capi.ptr(e, C.c_double),
                                       _capi.ptr(vecs, C.c_double)))
        order = np.argsort(dd)
        np.testing.assert_allclose(dd[order], np.linalg.eigvalsh(t), atol=1e-12 * max(1, np.abs(t).max()))
        np.testing.assert_allclose(t @ vecs, vecs * dd[None, :], atol=1e-11)
        np.testing.assert_allclose(vecs.T @ vecs, np.eye(n), atol=1e-11)


@pytest.mark.parametrize("seed", range(12))
def test_random_spectra_with_clusters_and_signs(host_lib, seed):
    """Matrices with a prescribed spectrum: a few leading eigenvalues (some negative, some only
    1e-3 apart) over a bulk; the pairs returned must be the ones of largest magnitude, in
    descending algebraic order, each with a small residual."""
    rng = np.random.default_rng(900 + seed)
    n = int(rng.integers(40, 260))
    k = int(rng.integers(1, 6))
    lead = rng.uniform(3.0, 9.0, size=6) * rng.choice([-1.0, 1.0], size=6, p=[0.3, 0.7])
    if seed % 3 == 0:
        lead[1] = lead[0] * (1 - 1e-3)                    # a tight pair at the top
    bulk = rng.uniform(-1.0, 1.0, size=n - 6)
    w = np.concatenate([lead, bulk])
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    a = (q * w) @ q.T
    a = (a + a.T) / 2
    evals, evecs, info = run(host_lib, a, k)
    assert info[3] == 1 and info[0] == k
    want = np.sort(w[np.argsort(-np.abs(w))[:k]])[::-1]
    np.testing.assert_allclose(evals, want, rtol=0, atol=1e-10)
    for lam, v in zip(evals, evecs):
        assert np.linalg.norm(a @ v - lam * v) < 1e-9
    assert info[4] <= info[1]                             # fewer host round trips than Lanczos steps
