"""CPU: the certified substitution sweep mode (csrc/certify.cuh, fast_kernels.cuh), executed through the host build of
the very templates the kernels instantiate (tests/host_shim), against the rounding-exact oracle (oracle/sdc_exact.c,
itself pinned by the reference's golden vectors).

Checked: (1) niter / converged / err of EVERY env bit-equal to the oracle (certified envs by the margin argument,
fallback envs because the exact kernel body re-ran them); (2) fallback envs bit-equal in u, r, ||r||; certified envs
within 1e-12 relative (the north star's tolerance); (3) the PROOF OBLIGATION itself, sweep by sweep: the margin the
kernel carries dominates the actual | ||r~_k|| - ||r^ref_k|| | (reference residuals of every sweep from the oracle's
collect_states buffer, sdc_env.py:239-240)."""
import numpy as np
import pytest

from oracle import exact
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner
from tests import host_shim

RTOL = 1e-12  # BASELINE.json north_star: residuals / states within 1e-12 relative


def _workload(kind, M, n, seed, cplx=False):
    rng = np.random.default_rng(seed)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    if kind == "uniform":
        act = rng.uniform(-1, 1, (n, M))
    else:  # near the MIN preconditioner: about half of the envs converge
        x = np.diag(fixed_preconditioner("min", M, collocation_matrix(M)))
        if not x.any():  # M without a tabulated MIN diagonal: the LU diagonal is a decent preconditioner too
            x = np.diag(fixed_preconditioner("LU", M, collocation_matrix(M)))
        act = 2 * (x[None] + rng.uniform(-0.02, 0.02, (n, M))) - 1
    if cplx:
        act = (0.5 * (act + 1)) + 1j * rng.uniform(-0.02, 0.02, (n, M))
    return lam, act


def _run(M, lam, act, *, variant=0, cplx=False, prec=None, strategy="iteration_only", do_scale=True, restol=1e-10,
         run_fallback=True):
    n = lam.shape[0]
    Q = collocation_matrix(M)
    d = host_shim.make_desc("sdc-v0", M, variant=variant, cplx=cplx, prec=prec, strategy=strategy,
                            do_scale=do_scale and not cplx, restol=restol)
    d.sweep_mode = 1
    b = host_shim.ShimBatch(d, n, entry="shim_step_certified")
    b.run_fallback = run_fallback
    b.reset(lam)
    out = b.step(None if prec is not None else act)
    u, r = exact.reset(Q, 1.0, lam, variant=variant)
    niter = np.zeros(n, np.int32)
    col = np.zeros((n, 2 * M, 50), np.complex128)
    ref = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), act,
                     prec_type="fixed" if prec is not None else "diag",
                     Qd_fixed=None if prec is None else fixed_preconditioner(prec, M, Q), do_scale=do_scale and not cplx,
                     reward_strategy=strategy, restol=restol, variant=variant, collect_states=col)
    fb = np.zeros(n, bool)
    fb[b.fallback_list[: b.fallback_count[0]]] = True
    return b, out, dict(u=u, r=r, niter=niter, col=col, **ref), fb


def _check(M, out, ref, fb, run_fallback=True):
    sel = slice(None) if run_fallback else ~fb
    assert np.array_equal(out["niter"][sel], ref["niter"][sel]), "iteration counts"
    assert np.array_equal(out["conv"][sel], ref["done"][sel]), "converged flags"
    assert np.array_equal(out["err"][sel], ref["err"][sel]), "err flags"
    if run_fallback and fb.any():  # re-run by the exact kernel body: every bit
        assert np.array_equal(out["term_u"][fb], ref["u"][fb]) and np.array_equal(out["term_r"][fb], ref["r"][fb])
        assert np.array_equal(out["residual"][fb], ref["resnorm"][fb])
    c = ~fb
    scale = 1.0 + np.abs(ref["u"][c]).max(axis=1, keepdims=True)
    assert np.all(np.abs(out["term_u"][c] - ref["u"][c]) <= RTOL * scale), "u of certified envs"
    # r = u0 - C u: rounding-level differences scale with ||u0|| + ||C|| ||u||, not with the (converged) residual
    cn = 1.0 + 100.0 * np.abs(ref["u"][c]).max(axis=1, keepdims=True)
    assert np.all(np.abs(out["term_r"][c] - ref["r"][c]) <= RTOL * cn), "r of certified envs"
    ok = np.abs(out["residual"][c] - ref["resnorm"][c]) <= RTOL * cn[:, 0]
    assert np.all(ok), "||r|| of certified envs"


def _margin_dominates(M, b, ref, fb):
    """| ||r~_k|| - ||r^ref_k|| | <= margin_k for every sweep the substitution kernel ran (also on envs that were
    later handed to the exact kernel: the bound must hold up to the sweep at which they became ambiguous)."""
    tr, col, niter = b.trace, ref["col"], ref["niter"]
    worst, checked = 0.0, 0
    for i in range(tr.shape[0]):
        for k in range(1, min(int(niter[i]), 49) + 1):  # the reference buffer holds sweeps 1..49
            m = tr[i, k - 1, 2 * M + 1]
            if np.isnan(m):
                break
            diff = abs(np.abs(col[i, M:, k]).max() - tr[i, k - 1, 2 * M])
            cdiff = np.abs(col[i, M:, k] - (tr[i, k - 1, 0:2 * M:2] + 1j * tr[i, k - 1, 1:2 * M:2])).max()
            assert diff <= m and cdiff <= m, (i, k, diff, cdiff, m)
            worst = max(worst, cdiff / m)
            checked += 1
    return worst, checked


@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("kind", ["uniform", "good"])
def test_certified_equals_oracle(M, kind):
    n = 600 if M <= 5 else 300
    lam, act = _workload(kind, M, n, seed=10 * M + (kind == "good"))
    b, out, ref, fb = _run(M, lam, act)
    _check(M, out, ref, fb)
    worst, checked = _margin_dominates(M, b, ref, fb)
    assert checked > n and worst < 1.0
    if kind == "uniform" and M >= 3:
        assert fb.mean() < 0.05  # the benchmark workload is certified almost entirely


def test_certified_envs_alone_are_right_without_the_fallback():
    """Leave the fallback list unprocessed: every env that is NOT on it must already carry the reference's decisions,
    and the envs on it must be untouched (the exact kernel needs their original state)."""
    M, n = 5, 1500
    lam, act = _workload("good", M, n, seed=77)
    b, out, ref, fb = _run(M, lam, act, run_fallback=False)
    assert 0 < fb.sum() < n // 2
    _check(M, out, ref, fb, run_fallback=False)
    u0, r0 = exact.reset(collocation_matrix(M), 1.0, lam)
    assert np.array_equal(out["u"][fb], u0[fb]) and np.array_equal(out["r"][fb], r0[fb])


@pytest.mark.parametrize("variant", [0, 1])
def test_certified_blas_variants_and_complex_actions(variant):
    M, n = 5, 500
    lam, act = _workload("good", M, n, seed=5, cplx=True)
    b, out, ref, fb = _run(M, lam, act, variant=variant, cplx=True)
    _check(M, out, ref, fb)
    assert _margin_dominates(M, b, ref, fb)[0] < 1.0


@pytest.mark.parametrize("prec", ["min", "zeros"])
def test_certified_fixed_diagonal_preconditioners(prec):
    M, n = 5, 500
    lam, act = _workload("uniform", M, n, seed=3)
    b, out, ref, fb = _run(M, lam, act, prec=prec)
    _check(M, out, ref, fb)
    assert _margin_dominates(M, b, ref, fb)[0] < 1.0


@pytest.mark.parametrize("strategy", ["residual_change", "gauss_kernel", "fast_convergence", "smooth_fast_convergence",
                                      "smoother_fast_convergence"])
def test_certified_rewards_within_tolerance(strategy):
    M, n = 5, 400
    lam, act = _workload("good", M, n, seed=11)
    b, out, ref, fb = _run(M, lam, act, strategy=strategy)
    _check(M, out, ref, fb)
    a, r = out["reward"], ref["reward"]
    if strategy == "gauss_kernel":
        # exp(-(nr/restol)^2 / 2) amplifies a relative change of nr by (nr/restol)^2: compare where that is moderate
        # (x = nr/restol <= 3: a 1e-4 relative difference of nr becomes <= 1e-3 of the reward)
        sel = ~ref["err"] & (ref["resnorm"] < 3e-10)
        assert sel.any() and np.all(np.abs(a[sel] - r[sel]) <= 5e-3 * np.abs(r[sel]))
    else:
        # these rewards are functions of log ||r||: near convergence ||r|| ~ 1e-10 carries the ABSOLUTE rounding-level
        # difference eps ||C|| ||u|| ~ 1e-14 (any rounding sequence other than the reference's does), i.e. 1e-4
        # relative, which log() turns into ~1e-5 of the reward
        assert np.all(np.abs(a - r) <= 3e-5 * np.maximum(1.0, np.abs(r))), np.abs(a - r).max()
    assert np.array_equal(a[fb], r[fb]) or np.allclose(a[fb], r[fb], rtol=1e-14, atol=0)


def test_certified_loose_and_tight_tolerances():
    M, n = 5, 500
    lam, act = _workload("good", M, n, seed=21)
    for restol in (1e-6, 1e-13):
        b, out, ref, fb = _run(M, lam, act, restol=restol)
        _check(M, out, ref, fb)


def test_certified_real_lambda_and_positive_real_part():
    """Im(lambda) = 0 (the reference's default interval) and unstable lambda (Re > 0, P can be near singular): the
    certificate must stay valid or hand over to the exact kernel."""
    M, n = 5, 600
    rng = np.random.default_rng(1)
    lam = np.concatenate([rng.uniform(-100, 0, n // 2) + 0j, rng.uniform(0, 5, n // 2) + 1j * rng.uniform(-3, 3, n // 2)])
    act = rng.uniform(-1, 1, (n, M))
    b, out, ref, fb = _run(M, lam, act)
    _check(M, out, ref, fb)
    assert _margin_dominates(M, b, ref, fb)[0] < 1.0
