"""GPU: sweep_mode='certified' (certificate kernel -> substitution sweeps -> exact kernel over the fallback list) through
the C ABI, against the exact mode of the same library (itself bit-equal to the reference's golden vectors) and against
the CPU oracle.  niter / done / converged / err: bit-equal for every env.  u, r, ||r||: within 1e-12 relative to
||u0|| + ||C|| ||u|| for certified envs, bit-equal for the envs the exact kernel re-ran.  The observation returned by an
auto-reset step (the next episode's initial state) is bit-equal always."""
import numpy as np
import pytest
import torch

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner

pytestmark = pytest.mark.gpu
RTOL = 1e-12
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)


def _actions(kind, M, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(-1, 1, (n, M))
    x = np.diag(fixed_preconditioner("min", M, collocation_matrix(M)))
    if not x.any():
        x = np.diag(fixed_preconditioner("LU", M, collocation_matrix(M)))
    return 2 * (x[None] + rng.uniform(-0.02, 0.02, (n, M))) - 1


def _pair(M, n, **kw):
    a = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=3, sweep_mode="exact", **KW, **kw)
    b = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=3, sweep_mode="certified", **KW, **kw)
    return a, b


def _compare_step(ea, eb, act, M):
    n = ea.num_envs
    t = torch.as_tensor(act, device=ea.device)
    oa = {k: v.clone() for k, v in ea.step_tensor(t).items()}
    ob = {k: v.clone() for k, v in eb.step_tensor(t).items()}
    nfb, _ = eb.fallback_stats()
    fb = np.zeros(n, bool)
    fb[eb.fallback_list[:nfb].cpu().numpy()] = True
    for k in ("niter", "flags"):
        assert torch.equal(oa[k], ob[k]), k
    assert torch.equal(oa["lam"], ob["lam"])
    ta, tb = oa["terminal"].cpu().numpy(), ob["terminal"].cpu().numpy()
    assert np.array_equal(ta[:, fb], tb[:, fb]), "fallback envs must be bit-equal"
    ua = np.abs(ta[: 2 * M]).max(axis=0)
    scale_u, scale_r = 1.0 + ua, 1.0 + 100.0 * ua
    assert np.all(np.abs(ta[: 2 * M] - tb[: 2 * M]) <= RTOL * scale_u)
    assert np.all(np.abs(ta[2 * M:] - tb[2 * M:]) <= RTOL * scale_r)
    ra, rb = oa["residual"].cpu().numpy(), ob["residual"].cpu().numpy()
    assert np.array_equal(ra[fb], rb[fb])
    assert np.all(np.abs(ra - rb) <= RTOL * scale_r)
    assert torch.equal(oa["reward"], ob["reward"])  # iteration_only: a function of niter / err alone
    return fb


@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("kind", ["uniform", "good"])
def test_certified_matches_exact_mode_over_autoreset_steps(M, kind):
    n = 20000
    ea, eb = _pair(M, n)
    ea.reset()
    eb.reset()
    total_fb = 0
    for s in range(3):
        act = _actions(kind, M, n, seed=100 * M + s)
        fb = _compare_step(ea, eb, act, M)
        total_fb += int(fb.sum())
        # the returned observation is the next episode's initial state: exact arithmetic in both modes
        assert torch.equal(ea.S, eb.S) and torch.equal(ea.lam, eb.lam) and torch.equal(ea.resnorm, eb.resnorm)
        assert torch.equal(ea.episodes, eb.episodes) and torch.equal(ea.rng_ctr, eb.rng_ctr)
    assert eb.fallback_stats()[1] == total_fb
    if kind == "uniform" and M >= 3:
        assert total_fb < 0.05 * 3 * n


def test_certified_against_cpu_oracle():
    M, n = 5, 20000
    Q = collocation_matrix(M)
    rng = np.random.default_rng(0)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    act = _actions("good", M, n, seed=1)
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, sweep_mode="certified", autoreset=False, **KW)
    env.reset(lam=lam)
    _, rew, done, infos = env.step(act)
    u, r = exact.reset(Q, 1.0, lam)
    niter = np.zeros(n, np.int32)
    out = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), act)
    assert np.array_equal(infos.niter, niter)
    assert np.array_equal((infos.flags & 2) != 0, out["done"]) and np.array_equal((infos.flags & 4) != 0, out["err"])
    snap = env._snapshot()
    scale = 1.0 + 100.0 * np.abs(u).max(axis=1)
    assert np.all(np.abs(snap["obs"][:, 0] - u) <= RTOL * scale[:, None])
    assert np.all(np.abs(snap["obs"][:, 1] - r) <= RTOL * scale[:, None])
    assert np.all(np.abs(infos.residual - out["resnorm"]) <= RTOL * scale)
    nfb, _ = env.fallback_stats()
    assert 0 < nfb < 0.3 * n
    # a second step would start from the approximate state of the first one: refused
    with pytest.raises(_lib.SdcGymError):
        env.step(act)
    env.reset(lam=lam)
    env.step(act)


@pytest.mark.parametrize("kw", [dict(prec="min"), dict(free_action_space=True, do_scale=False),
                                dict(reward_iteration_only=False), dict(blas_variant=_lib.BLAS_HASWELL)])
def test_certified_configurations(kw):
    M, n = 5, 8192
    base = {k: v for k, v in KW.items() if k not in kw}
    a = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=9, sweep_mode="exact", **base, **kw)
    b = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=9, sweep_mode="certified", **base, **kw)
    a.reset()
    b.reset()
    act = _actions("good", M, n, seed=4)
    if kw.get("free_action_space"):
        act = 0.5 * (act + 1) + 1j * np.random.default_rng(2).uniform(-0.02, 0.02, act.shape)
    t = None if "prec" in kw else torch.as_tensor(act, device=a.device)
    oa = {k: v.clone() for k, v in a.step_tensor(t).items()}
    ob = {k: v.clone() for k, v in b.step_tensor(t).items()}
    assert torch.equal(oa["niter"], ob["niter"]) and torch.equal(oa["flags"], ob["flags"])
    assert torch.equal(a.S, b.S)
    if kw.get("reward_iteration_only") is False:
        assert torch.allclose(oa["reward"], ob["reward"], rtol=3e-5, atol=3e-5)
    else:
        assert torch.equal(oa["reward"], ob["reward"])


def test_certified_host_step_through_the_result_block():
    """The numpy drop-in path (sdcgym_pipe_step_block, chunked) in certified mode: same outputs as the exact mode."""
    M, n = 5, 150000  # large enough for the chunked pipeline
    ea, eb = _pair(M, n)
    ea.reset()
    eb.reset()
    for s in range(2):
        act = _actions("uniform" if s else "good", M, n, seed=50 + s)
        oa, ra, da, ia = ea.step(act)
        ob, rb, db, ib = eb.step(act)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db)
        assert np.array_equal(ia.niter, ib.niter) and np.array_equal(ia.flags, ib.flags)
        assert np.array_equal(ia.lam, ib.lam)
        assert np.all(np.abs(ia.residual - ib.residual) <= 1e-10 * (1 + np.abs(ia.residual)))
    assert eb.fallback_stats()[1] > 0


def test_certified_falls_back_to_exact_for_unsupported_combinations():
    """sdc-v1, collect_states and dense Q_delta have no certificate: the exact kernels run, results bit-equal."""
    M, n = 5, 4096
    for kw in (dict(envname="sdc-v1"), dict(envname="sdc-v0", collect_states=True),
               dict(envname="sdc-v0", prec="LU")):
        name = kw.pop("envname")
        a = sdc_gym_b200.make(name, num_envs=n, M=M, seed=1, sweep_mode="exact", **KW, **kw)
        b = sdc_gym_b200.make(name, num_envs=n, M=M, seed=1, sweep_mode="certified", **KW, **kw)
        a.reset()
        b.reset()
        act = _actions("good", M, n, seed=2)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(ia.residual, ib.residual)
