"""Gauss-Radau-right collocation on [tleft, tright], restated from the published definition.

The reference takes its collocation matrix from the un-vendored third-party package pySDC
(``CollGaussRadau_Right(M, 0, 1).Qmat[1:, 1:]``, reference call sites ``sdc_gym/envs/sdc_env.py:53-54``,
``dp_playground.py:79-80,188-189``; pySDC version unpinned in ``setup.py:8``).  pySDC is not installed
here and no reference test pins the bits of ``Q``, so this module restates the mathematical definition

    nodes  tau_0 < ... < tau_{M-1} = 1 : right Radau points, i.e. the roots of P_{M-1}(x) - P_M(x)... on [-1, 1]
                                          mapped to [0, 1] (equivalently: roots of the Jacobi polynomial
                                          P^{(1,0)}_{M-1} plus the right end point)
    Q[m, j] = int_0^{tau_m} l_j(s) ds   : l_j = Lagrange basis polynomial on the nodes

and evaluates it in 80-digit ``decimal`` arithmetic, rounding once to binary64 at the end.  Every entry is
therefore the correctly rounded double of the exact value: deterministic on every host, no BLAS, no libm.
``Q`` is an explicit input of the oracle and of the CUDA kernels alike (parity never depends on how it was
produced).

Sanity identities used by the tests: ``Q @ 1 = tau``, ``Q @ tau = tau**2 / 2``, ``Q[M-1, M-1] = 1 / M**2``.
"""
from __future__ import annotations

import functools
from decimal import Decimal, getcontext, localcontext

import numpy as np

_PREC = 80


def _legendre(n: int, x: Decimal):
    """Return (P_n(x), P_{n-1}(x)) by the three-term recurrence."""
    p0, p1 = Decimal(1), x
    if n == 0:
        return p0, Decimal(0)
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    return p1, p0


def _legendre_deriv(n: int, x: Decimal, pn: Decimal, pnm1: Decimal) -> Decimal:
    return n * (x * pn - pnm1) / (x * x - 1)


def _newton_roots(f_and_df, guesses):
    roots = []
    for x in guesses:
        x = Decimal(x)
        for _ in range(200):
            f, df = f_and_df(x)
            # deflate already-found roots so every guess converges to a new one
            s = sum((1 / (x - r) for r in roots), Decimal(0))
            dx = f / (df - f * s)
            x -= dx
            if abs(dx) < Decimal(10) ** (-(_PREC - 10)):
                break
        roots.append(x)
    return sorted(roots)


def _gauss_legendre(n: int):
    """n-point Gauss-Legendre nodes/weights on [-1, 1] in Decimal."""
    import math

    def f_and_df(x):
        pn, pnm1 = _legendre(n, x)
        return pn, _legendre_deriv(n, x, pn, pnm1)

    guesses = [repr(math.cos(math.pi * (i + 0.75) / (n + 0.5))) for i in range(n)]
    xs = _newton_roots(f_and_df, guesses)
    ws = []
    for x in xs:
        pn, pnm1 = _legendre(n, x)
        d = _legendre_deriv(n, x, pn, pnm1)
        ws.append(2 / ((1 - x * x) * d * d))
    return xs, ws


def _radau_right_nodes_pm1(M: int):
    """Right Radau nodes on [-1, 1]: the M-1 interior roots of (P_{M-1}(x) - P_M(x)) / (1 - x), plus x = 1.

    (For the *left* Radau family the polynomial is P_{M-1} + P_M; mirroring x -> -x gives the right one.)
    """
    import math

    if M == 1:
        return [Decimal(1)]

    def f_and_df(x):
        pM, pMm1 = _legendre(M, x)
        pMm1_, pMm2 = _legendre(M - 1, x)
        g = pMm1 - pM
        dg = _legendre_deriv(M - 1, x, pMm1_, pMm2) - _legendre_deriv(M, x, pM, pMm1)
        # divide out the known root at x = 1:  h = g / (1 - x),  h' = (dg (1-x) + g) / (1-x)^2
        omx = 1 - x
        return g / omx, (dg * omx + g) / (omx * omx)

    # Chebyshev-like interior guesses, strictly inside (-1, 1)
    guesses = [repr(-math.cos(math.pi * (2 * i + 1) / (2 * M - 1))) for i in range(M - 1)]
    roots = _newton_roots(f_and_df, guesses)
    return roots + [Decimal(1)]


@functools.lru_cache(maxsize=None)
def _radau_right_decimal(M: int, tleft: str, tright: str):
    with localcontext() as ctx:
        ctx.prec = _PREC
        a, b = Decimal(tleft), Decimal(tright)
        half = Decimal(1) / 2
        nodes = [a + (b - a) * (x + 1) * half for x in _radau_right_nodes_pm1(M)]
        gx, gw = _gauss_legendre(max(M, 1))

        def lagrange(j, s):
            v = Decimal(1)
            for k in range(M):
                if k != j:
                    v *= (s - nodes[k]) / (nodes[j] - nodes[k])
            return v

        def integrate(j, lo, hi):
            h = (hi - lo) * half
            return h * sum((w * lagrange(j, lo + h * (x + 1)) for x, w in zip(gx, gw)), Decimal(0))

        Q = [[integrate(j, a, nodes[m]) for j in range(M)] for m in range(M)]
        weights = [integrate(j, a, b) for j in range(M)]
        delta = [nodes[0] - a] + [nodes[m] - nodes[m - 1] for m in range(1, M)]
        to_f = lambda v: float(v)  # noqa: E731  (correct rounding Decimal -> binary64)
        return (
            tuple(to_f(v) for v in nodes),
            tuple(tuple(to_f(v) for v in row) for row in Q),
            tuple(to_f(v) for v in weights),
            tuple(to_f(v) for v in delta),
        )


class CollGaussRadauRight:
    """Stand-in for pySDC's ``CollGaussRadau_Right(num_nodes, tleft, tright)``.

    Exposes the three attributes the reference touches (``sdc_env.py:53-54,186``): ``Qmat`` of shape
    (M+1, M+1) with a zero first row and column, ``delta_m`` and ``num_nodes`` (plus ``nodes``, ``weights``).
    """

    def __init__(self, num_nodes: int, tleft: float = 0, tright: float = 1):
        if num_nodes < 1:
            raise ValueError("need at least one collocation node")
        nodes, Q, weights, delta = _radau_right_decimal(int(num_nodes), repr(float(tleft)), repr(float(tright)))
        self.num_nodes = int(num_nodes)
        self.tleft = tleft
        self.tright = tright
        self.nodes = np.array(nodes, dtype=np.float64)
        self.weights = np.array(weights, dtype=np.float64)
        self.delta_m = np.array(delta, dtype=np.float64)
        self.Qmat = np.zeros((self.num_nodes + 1, self.num_nodes + 1), dtype=np.float64)
        self.Qmat[1:, 1:] = np.array(Q, dtype=np.float64)


def collocation_matrix(M: int) -> np.ndarray:
    """``CollGaussRadau_Right(M, 0, 1).Qmat[1:, 1:]`` as a fresh C-contiguous (M, M) float64 array."""
    return np.ascontiguousarray(CollGaussRadauRight(M, 0, 1).Qmat[1:, 1:]).copy()
