"""Fixed preconditioners Q_delta selected by the reference's ``prec`` constructor argument
(``sdc_gym/envs/sdc_env.py:134-191``) and the action layouts of ``dp_playground.py:194-207``.

These are one-off host-side constants (an M x M real matrix handed to the kernels as data); nothing here
is on the per-step path.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .collocation import CollGaussRadauRight

# hard-coded "MIN" diagonals, reference sdc_env.py:147-177
_MIN_DIAGONALS = {
    7: [0.15223871397682717, 0.12625448001038536, 0.08210714764924298, 0.03994434742760019,
        0.1052662547386142, 0.14075805578834127, 0.15636085758812895],
    5: [0.2818591930905709, 0.2011358490453793, 0.06274536689514164, 0.11790265267514095, 0.1571629578515223],
    4: [0.3198786751412953, 0.08887606314792469, 0.1812366328324738, 0.23273925017954],
    3: [0.3203856825077055, 0.1399680686269595, 0.3716708461097372],
}

PREC_TYPES = ("diag", "lower_diag", "lower_tri", "strictly_lower_tri")


def fixed_preconditioner(prec: str, M: int, Q: np.ndarray | None = None) -> np.ndarray:
    """Return the real (M, M) matrix the reference builds for ``prec`` in {'LU','min','EE','zeros'}."""
    coll = CollGaussRadauRight(M, 0, 1)
    if Q is None:
        Q = coll.Qmat[1:, 1:]
    Q = np.asarray(Q, dtype=np.float64)
    if prec.upper() == "LU":
        # sdc_env.py:141-144:  U.T of the pivoted LU of Q.T.  The reference passes overwrite_a=True on a
        # non-contiguous view (scipy then works on an internal copy); a private copy without the flag is the
        # same LAPACK getrf on the same values and cannot clobber the caller's Q.
        _, _, U = scipy.linalg.lu(np.array(Q.T, dtype=np.float64, order="C", copy=True))
        return np.ascontiguousarray(U.T)
    if prec.lower() == "min":
        Qd = np.zeros_like(Q)
        np.fill_diagonal(Qd, _MIN_DIAGONALS.get(M, np.zeros(M)))
        return Qd
    if prec.upper() == "EE":
        Qd = np.zeros_like(Q)
        for m in range(M):
            Qd[m, 0:m] = coll.delta_m[1:m + 1]
        return Qd
    if prec.lower() == "zeros":
        return np.zeros_like(Q)
    raise NotImplementedError(f"unknown preconditioner {prec!r}")


def num_actions(M: int, prec_type: str) -> int:
    """Length of the action / network output for ``prec_type`` (dp_playground.py:194-207)."""
    return {"diag": M, "lower_diag": M - 1, "lower_tri": M * (M + 1) // 2,
            "strictly_lower_tri": M * (M - 1) // 2}[prec_type]


def qdmat_from_output(output, M: int, prec_type: str) -> np.ndarray:
    """Host restatement of ``SpectralRadiusLoss.get_qdmat`` (dp_playground.py:194-207) for one sample."""
    output = np.asarray(output)
    if prec_type == "diag":
        return np.diag(output)
    if prec_type == "lower_diag":
        return np.diag(output, k=-1)
    Qd = np.zeros((M, M), dtype=output.dtype)
    if prec_type == "lower_tri":
        Qd[np.tril_indices(M)] = output
    elif prec_type == "strictly_lower_tri":
        Qd[np.tril_indices(M, k=-1)] = output
    else:
        raise ValueError(f"unknown prec_type {prec_type!r}")
    return Qd
