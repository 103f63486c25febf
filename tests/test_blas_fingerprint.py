"""CPU: the oracle's primitives against the numpy/OpenBLAS of THIS host (SURVEY Appendix B "BLAS fingerprint").

If the host dispatches OpenBLAS to a different core than the golden vectors were made with, the matching
variant is selected; the test then pins that the variant flag tracks the live library.
"""
import numpy as np
import pytest

from oracle import exact
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.vec_env import detect_blas_variant
from tests.helpers import assert_same

VARIANT = detect_blas_variant()


def test_complex_abs_matches_numpy():
    rng = np.random.default_rng(0)
    v = (rng.uniform(-1, 1, 4000) + 1j * rng.uniform(-1, 1, 4000)) * 10.0 ** rng.uniform(-12, 3, 4000)
    got = np.array([exact.cabs(z) for z in v])
    assert_same(got, np.abs(v), "abs")
    for z in (0j, 1e-320 + 0j, complex(np.inf, 1), complex(1, -np.inf), complex(np.nan, 1), complex(np.nan, np.inf)):
        a, b = exact.cabs(z), np.abs(np.array([z]))[0]
        assert (a == b) or (np.isnan(a) and np.isnan(b)), (z, a, b)


@pytest.mark.parametrize("M", range(2, 10))
def test_zgemv_matches_numpy(M):
    rng = np.random.default_rng(M)
    for _ in range(100):
        A = (rng.uniform(-1, 1, (M, M)) + 1j * rng.uniform(-1, 1, (M, M))) * 10.0 ** rng.uniform(-3, 3)
        x = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
        assert_same(exact.zgemv(A, x, VARIANT), A @ x, f"zgemv M={M}")


@pytest.mark.parametrize("M", range(2, 10))
def test_inverse_matches_numpy(M):
    rng = np.random.default_rng(100 + M)
    Q = collocation_matrix(M)
    for _ in range(60):
        A = rng.uniform(-1, 1, (M, M)) + 1j * rng.uniform(-1, 1, (M, M))
        assert_same(exact.cinv(A, VARIANT), np.linalg.inv(A), f"inv dense M={M}")
        z = complex(rng.uniform(-100, 0), rng.uniform(-10, 0))
        P = np.eye(M) - z * np.tril(rng.uniform(0, 1, (M, M)))
        assert_same(exact.cinv(P, VARIANT), np.linalg.inv(P), f"inv lower-tri M={M}")
        P = np.eye(M) - z * np.diag(rng.uniform(0, 1, M))
        assert_same(exact.cinv(P, VARIANT), np.linalg.inv(P), f"inv diag M={M}")
        P = np.eye(M) - z * np.tril(Q)
        assert_same(exact.cinv(P, VARIANT), np.linalg.inv(P), f"inv tril(Q) M={M}")


def test_scalar_times_matrix_matches_numpy():
    rng = np.random.default_rng(7)
    for M in (3, 5, 7, 9):
        Q = collocation_matrix(M)
        for _ in range(50):
            lam = complex(rng.uniform(-100, 0), rng.uniform(-10, 0))
            u, r = exact.reset(Q, 1.0, [lam], VARIANT)
            C = np.eye(M) - lam * 1.0 * Q
            assert_same(r[0], np.ones(M, np.complex128) - C @ np.ones(M, np.complex128), "reset residual")
