"""TEST INFRASTRUCTURE ONLY - ctypes front end of ``oracle/sdc_exact.c`` (the rounding-exact CPU oracle).

Importable only from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg.  Nothing under ``sdc_gym_b200/`` imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsdc_oracle.so")

PREC_TYPES = {"diag": 0, "lower_diag": 1, "lower_tri": 2, "strictly_lower_tri": 3, "fixed": 4}
REWARD_STRATEGIES = {
    "iteration_only": 0,
    "residual_change": 1,
    "gauss_kernel": 2,
    "fast_convergence": 3,
    "smooth_fast_convergence": 4,
    "smoother_fast_convergence": 5,
    "spectral_radius": 6,
}
VARIANT_SKYLAKEX, VARIANT_HASWELL = 0, 1

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_bp = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no-op when the .so is newer than the source)."""
    src = os.path.join(_HERE, "sdc_exact.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.sdc_oracle_cabs.restype = ctypes.c_double
        L.sdc_oracle_cabs.argtypes = [ctypes.c_double, ctypes.c_double]
        L.sdc_oracle_num_actions.restype = ctypes.c_int
        L.sdc_oracle_zgemv.argtypes = [ctypes.c_int, _dp, _dp, _dp, ctypes.c_int]
        L.sdc_oracle_cinv.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int]
        L.sdc_oracle_reset.argtypes = [ctypes.c_int, _dp, ctypes.c_double, ctypes.c_int64, _dp, _dp, _dp, ctypes.c_int]
        step_args = [
            ctypes.c_int, _dp, ctypes.c_double, ctypes.c_int64, ctypes.c_int, _dp,  # M Q dt N prec Qd_fixed
            _dp, ctypes.c_int, ctypes.c_int, _dp,  # action is_complex do_scale lam
            _dp, _dp, _ip, _dp,  # u r niter rinit
            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            _dp, _bp, _dp, _bp, ctypes.c_int, _dp,
        ]
        L.sdc_oracle_step_v1.argtypes = step_args
        L.sdc_oracle_step_v0.argtypes = step_args
        for f in (L.sdc_oracle_zgemv, L.sdc_oracle_cinv, L.sdc_oracle_reset, L.sdc_oracle_step_v0, L.sdc_oracle_step_v1):
            f.restype = None
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def cabs(z) -> float:
    return lib().sdc_oracle_cabs(float(np.real(z)), float(np.imag(z)))


def zgemv(A, x, variant=0):
    A = np.ascontiguousarray(A, np.complex128)
    x = np.ascontiguousarray(x, np.complex128)
    y = np.empty(A.shape[0], np.complex128)
    lib().sdc_oracle_zgemv(A.shape[0], _d(A), _d(x), _d(y), variant)
    return y


def cinv(P, variant=0):
    P = np.ascontiguousarray(P, np.complex128)
    out = np.empty_like(P)
    lib().sdc_oracle_cinv(P.shape[0], _d(P), _d(out), variant)
    return out


def num_actions(M, prec_type):
    return lib().sdc_oracle_num_actions(M, PREC_TYPES[prec_type])


def reset(Q, dt, lam, variant=0):
    """(u, r) of shape (N, M) complex128 for the lambdas ``lam`` (N,)."""
    Q = np.ascontiguousarray(Q, np.float64)
    lam = np.ascontiguousarray(lam, np.complex128).reshape(-1)
    M, N = Q.shape[0], lam.shape[0]
    u = np.empty((N, M), np.complex128)
    r = np.empty((N, M), np.complex128)
    lib().sdc_oracle_reset(M, _d(Q), float(dt), N, _d(lam), _d(u), _d(r), variant)
    return u, r


def step(kind, Q, dt, lam, u, r, niter, rinit, action, *, prec_type="diag", Qd_fixed=None, do_scale=True,
         reward_strategy="iteration_only", step_penalty=0.1, residual_weight=0.5, norm_factor=1.0,
         restol=1e-10, max_iters=50, variant=0, collect_states=None, want_pinv=False, use_doubles=True):
    """Advance N envs in place (u, r, niter are modified).  Returns dict(reward, done, resnorm, err[, pinv]).

    ``kind`` is 'sdc-v0' or 'sdc-v1'.  For 'sdc-v0', ``done`` is the *converged* flag (the env itself always
    reports done=True, ``sdc_env.py:259``).  ``use_doubles=False``: float32 / complex64 action space, i.e. the
    learned Q_delta entries are rounded to float32 (``sdc_env.py:100,109,138-140``).
    """
    Q = np.ascontiguousarray(Q, np.float64)
    M = Q.shape[0]
    lam = np.ascontiguousarray(lam, np.complex128).reshape(-1)
    N = lam.shape[0]
    assert u.dtype == np.complex128 and u.shape == (N, M) and u.flags.c_contiguous
    assert r.dtype == np.complex128 and r.shape == (N, M) and r.flags.c_contiguous
    assert niter.dtype == np.int32 and niter.shape == (N,)
    rinit = np.ascontiguousarray(rinit, np.complex128)
    pt = PREC_TYPES[prec_type]
    A = lib().sdc_oracle_num_actions(M, pt)
    is_c = 0
    if pt == 4:
        Qd_fixed = np.ascontiguousarray(Qd_fixed, np.float64)
        act = None
    else:
        action = np.asarray(action)
        is_c = int(np.iscomplexobj(action))
        act = np.ascontiguousarray(action, np.complex128 if is_c else np.float64).reshape(N, A)
    reward = np.empty(N, np.float64)
    done = np.empty(N, np.uint8)
    resnorm = np.empty(N, np.float64)
    err = np.empty(N, np.uint8)
    extra = None
    if kind == "sdc-v1":
        fn = lib().sdc_oracle_step_v1
        if want_pinv:
            extra = np.empty((N, M, M), np.complex128)
    else:
        fn = lib().sdc_oracle_step_v0
        extra = collect_states
        if extra is not None:
            assert extra.dtype == np.complex128 and extra.shape == (N, 2 * M, max_iters) and extra.flags.c_contiguous
    fn(M, _d(Q), float(dt), N, pt, _d(Qd_fixed), _d(act), is_c, int(bool(do_scale)) | (0 if use_doubles else 2), _d(lam),
       _d(u), _d(r), niter.ctypes.data_as(_ip), _d(rinit),
       REWARD_STRATEGIES[reward_strategy], float(step_penalty), float(residual_weight), float(norm_factor),
       float(restol), int(max_iters),
       _d(reward), done.ctypes.data_as(_bp), _d(resnorm), err.ctypes.data_as(_bp), int(variant), _d(extra))
    out = dict(reward=reward, done=done.astype(bool), resnorm=resnorm, err=err.astype(bool))
    if want_pinv and kind == "sdc-v1":
        out["pinv"] = extra
    return out
