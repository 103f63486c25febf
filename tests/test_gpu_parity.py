"""GPU: libsdcgym.so (through the C ABI, via SDCVecEnv) against the golden vectors of the reference env and
against the rounding-exact CPU oracle on seeded batches.  Iteration counts, flags, states and residual norms
must be bit-exact (value-exact: +0 == -0); rewards within 1e-14 relative (libm log/exp)."""
import numpy as np
import pytest

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner, num_actions
from tests.helpers import assert_reward_close, assert_same, case_ids
from tests.replay import replay_case

pytestmark = pytest.mark.gpu


class CudaBackend:
    def __init__(self, meta, g):
        self.env = sdc_gym_b200.make(
            meta["kind"], num_envs=meta["n"], M=meta["M"], dt=meta["dt"], restol=meta["restol"], prec=meta["prec"],
            prec_type=meta["prec_type"] if meta["prec"] is None else "diag", free_action_space=meta["cplx"],
            do_scale=meta["do_scale"], use_doubles=meta.get("use_doubles", True), reward_iteration_only=None,
            reward_strategy=meta["strategy"],
            step_penalty=meta["step_penalty"], residual_weight=meta["residual_weight"], norm_factor=meta["norm_factor"],
            collect_states=meta["collect"], Q=g["Q"], blas_variant=_lib.BLAS_SKYLAKEX, autoreset=False,
            lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
        self.n_act = 0 if meta["prec"] is not None else num_actions(meta["M"], meta["prec_type"])
        self.collect = meta["collect"]
        self.M = meta["M"]

    def _ur(self):
        snap = self.env._snapshot()
        return snap["obs"][:, 0].copy(), snap["obs"][:, 1].copy()

    def reset(self, lam):
        self.env.reset(lam=lam)
        return self._ur()

    def step(self, actions):
        if actions is None:
            actions = [np.empty(self.env.action_space.shape, dtype=self.env.action_space.dtype)
                       for _ in range(self.env.num_envs)]  # uninitialised, as the reference scripts pass them
        obs, rew, done, infos = self.env.step(actions)
        u, r = self._ur()
        f = infos.flags
        return dict(u=u, r=r, reward=rew, done=done, conv=(f & 2) != 0, err=(f & 4) != 0, niter=infos.niter,
                    residual=infos.residual)

    def old_states_host(self):
        import torch
        return torch.view_as_complex(self.env.old_states).cpu().numpy()


@pytest.mark.parametrize("name", case_ids())
def test_cuda_replays_golden(name):
    replay_case(name, CudaBackend)


# --------------------------------------------------------------------------------------------------------
# larger seeded batches against the C oracle
# --------------------------------------------------------------------------------------------------------
def _actions(rng, n, M, prec_type, mode, cplx):
    A = num_actions(M, prec_type)
    if cplx:
        return rng.uniform(0, 0.5, (n, A)) + 1j * rng.uniform(-0.1, 0.1, (n, A))
    if mode == "good" and prec_type == "diag":
        x = np.diag(fixed_preconditioner("min", M))
        return 2 * (x[None, :] + rng.uniform(-0.02, 0.02, (n, M))) - 1
    if prec_type == "diag":
        return rng.uniform(-1, 1, (n, A))
    return rng.uniform(0, 0.6, (n, A))


def _compare_batch(kind, M, n, *, prec=None, prec_type="diag", mode="uniform", cplx=False, strategy="iteration_only",
                   steps=1, seed=0):
    rng = np.random.default_rng(seed)
    Q = collocation_matrix(M)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    do_scale = (prec_type == "diag") and not cplx
    env = sdc_gym_b200.make(kind, num_envs=n, M=M, dt=1.0, restol=1e-10, prec=prec, prec_type=prec_type,
                            free_action_space=cplx, do_scale=do_scale, reward_iteration_only=None,
                            reward_strategy=strategy, blas_variant=_lib.BLAS_SKYLAKEX, autoreset=False,
                            lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    obs = env.reset(lam=lam)
    u, r = exact.reset(Q, 1.0, lam)
    assert_same(obs[:, 0], u, "reset u")
    assert_same(obs[:, 1], r, "reset r")
    rinit, niter = r.copy(), np.zeros(n, np.int32)
    Qd_fixed = fixed_preconditioner(prec, M, Q) if prec else None
    alive = np.ones(n, bool)
    hist = None
    for s in range(steps):
        act = None if prec else _actions(rng, n, M, prec_type, mode, cplx)
        obs, rew, done, infos = env.step(act if act is not None else np.zeros((n, M)))
        out = exact.step(kind, Q, 1.0, lam, u, r, niter, rinit, act, prec_type="fixed" if prec else prec_type,
                         Qd_fixed=Qd_fixed, do_scale=do_scale, reward_strategy=strategy)
        snap = env._snapshot()
        assert_same(snap["obs"][alive, 0], u[alive], f"step {s} u")
        assert_same(snap["obs"][alive, 1], r[alive], f"step {s} r")
        assert np.array_equal(infos.niter[alive], niter[alive]), f"step {s} niter"
        assert_same(infos.residual[alive], out["resnorm"][alive], f"step {s} residual")
        assert_reward_close(rew[alive], out["reward"][alive], f"step {s}")
        f = infos.flags
        assert np.array_equal(((f & 4) != 0)[alive], out["err"][alive]), f"step {s} err"
        if kind == "sdc-v1":
            assert np.array_equal(done[alive], out["done"][alive]), f"step {s} done"
            alive &= ~out["done"]
        else:
            assert np.array_equal(((f & 2) != 0)[alive], out["done"][alive]), f"step {s} converged"
            assert done.all()
        hist = np.bincount(infos.niter, minlength=51)
    return hist


@pytest.mark.parametrize("M", [3, 5, 7, 9])
@pytest.mark.parametrize("mode", ["uniform", "good"])
def test_v0_diag_batch_equals_oracle(M, mode):
    if mode == "good" and M == 9:
        pytest.skip("no MIN diagonal for M=9")
    hist = _compare_batch("sdc-v0", M, 20000, mode=mode, seed=M)
    if mode == "good":
        assert hist[:50].sum() > 0  # early exits are exercised


@pytest.mark.parametrize("M", [3, 5, 7, 9])
def test_v1_diag_rollout_equals_oracle(M):
    _compare_batch("sdc-v1", M, 4096, mode="good" if M != 9 else "uniform", steps=50, seed=10 + M,
                   strategy="residual_change")


@pytest.mark.parametrize("M", [3, 5, 7, 9])
@pytest.mark.parametrize("prec", ["LU", "min", "EE", "zeros"])
def test_v0_fixed_prec_batch_equals_oracle(M, prec):
    _compare_batch("sdc-v0", M, 4096, prec=prec, seed=20 + M)


@pytest.mark.parametrize("M", [3, 5, 7, 9])
@pytest.mark.parametrize("prec_type", ["lower_diag", "lower_tri", "strictly_lower_tri"])
def test_v0_learned_triangular_batch_equals_oracle(M, prec_type):
    _compare_batch("sdc-v0", M, 4096, prec_type=prec_type, seed=30 + M)
    _compare_batch("sdc-v0", M, 1024, prec_type=prec_type, cplx=True, seed=40 + M)


@pytest.mark.parametrize("prec_type", ["diag", "lower_tri"])
def test_v1_complex_actions_rollout(prec_type):
    _compare_batch("sdc-v1", 5, 2048, prec_type=prec_type, cplx=True, steps=12, seed=5)


def test_v1_fixed_LU_rollout():
    _compare_batch("sdc-v1", 5, 2048, prec="LU", steps=50, seed=6, strategy="residual_change")


def test_haswell_variant_equals_oracle_variant():
    """The second BLAS variant (unfused scalar tails) is checked against the oracle run with the same flag."""
    n, M = 8192, 5
    rng = np.random.default_rng(77)
    Q = collocation_matrix(M)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    for prec in (None, "LU"):
        env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, dt=1.0, restol=1e-10, prec=prec,
                                blas_variant=_lib.BLAS_HASWELL, autoreset=False)
        env.reset(lam=lam)
        u, r = exact.reset(Q, 1.0, lam, variant=1)
        act = 2 * (np.diag(fixed_preconditioner("min", M))[None] + rng.uniform(-0.02, 0.02, (n, M))) - 1
        _, _, _, infos = env.step(act)
        niter = np.zeros(n, np.int32)
        out = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), None if prec else act,
                         prec_type="fixed" if prec else "diag",
                         Qd_fixed=fixed_preconditioner(prec, M, Q) if prec else None, variant=1)
        snap = env._snapshot()
        assert_same(snap["obs"][:, 0], u); assert_same(snap["obs"][:, 1], r)
        assert np.array_equal(infos.niter, niter); assert_same(infos.residual, out["resnorm"])


@pytest.mark.parametrize("M", [2, 4, 6, 8])
def test_even_M_all_paths_equal_oracle(M):
    """collocation sizes without golden vectors: every kernel family against the oracle"""
    _compare_batch("sdc-v0", M, 3000, mode="uniform", seed=60 + M)
    _compare_batch("sdc-v0", M, 1500, prec="LU", seed=61 + M)
    _compare_batch("sdc-v0", M, 1500, prec_type="lower_tri", seed=62 + M)
    _compare_batch("sdc-v1", M, 1000, mode="uniform", steps=20, seed=63 + M, strategy="residual_change")
    _compare_batch("sdc-v1", M, 500, prec_type="strictly_lower_tri", cplx=True, steps=8, seed=64 + M)
