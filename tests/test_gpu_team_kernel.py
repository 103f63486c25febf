"""GPU: the lane-team step kernel (csrc/team_kernels.cuh, dense Q_delta at M >= 8) against the rounding-exact oracle:
ragged batch sizes (partial warps / partial teams), both BLAS variants, the fused auto-reset of sdc-v0 and the
per-env auto-reset of an sdc-v1 rollout, scaled residual rewards, and padding canaries around the state planes."""
import numpy as np
import pytest

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib, rng as host_rng
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner, num_actions
from tests.helpers import assert_reward_close, assert_same
from tests.test_gpu_parity import _compare_batch

pytestmark = pytest.mark.gpu

KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)


@pytest.mark.parametrize("M", [8, 9])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 13, 47, 1000, 4099])
def test_ragged_sizes_equal_oracle(M, n):
    """32 // M envs per warp, 4 warps per block: sizes that end inside a team, a warp and a block"""
    _compare_batch("sdc-v0", M, n, prec_type="lower_tri", seed=100 + n)


@pytest.mark.parametrize("M", [8, 9])
@pytest.mark.parametrize("prec_type", ["lower_diag", "strictly_lower_tri", "lower_tri"])
def test_v1_dense_rollout_equals_oracle(M, prec_type):
    _compare_batch("sdc-v1", M, 1500, prec_type=prec_type, steps=30, seed=7 + M, strategy="residual_change")
    _compare_batch("sdc-v1", M, 700, prec_type=prec_type, cplx=True, steps=10, seed=8 + M)


@pytest.mark.parametrize("M", [8, 9])
@pytest.mark.parametrize("prec", ["LU", "EE"])
def test_haswell_variant_dense(M, prec):
    """second BLAS variant (unfused scalar tails) through the team kernel"""
    n = 3000
    rng = np.random.default_rng(M)
    Q = collocation_matrix(M)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, prec=prec, autoreset=False,
                            **{**KW, "blas_variant": _lib.BLAS_HASWELL})
    env.reset(lam=lam)
    u, r = exact.reset(Q, 1.0, lam, variant=1)
    _, _, _, infos = env.step(np.zeros((n, M)))
    niter = np.zeros(n, np.int32)
    out = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), None, prec_type="fixed",
                     Qd_fixed=fixed_preconditioner(prec, M, Q), variant=1)
    snap = env._snapshot()
    assert_same(snap["obs"][:, 0], u); assert_same(snap["obs"][:, 1], r)
    assert np.array_equal(infos.niter, niter); assert_same(infos.residual, out["resnorm"])


@pytest.mark.parametrize("M", [8, 9])
def test_v0_autoreset_dense(M):
    """fused DummyVecEnv auto-reset: terminal observation = the solve, returned observation = reset state of draw 1"""
    n = 1234
    rng = np.random.default_rng(3)
    Q = collocation_matrix(M)
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=5, prec_type="lower_tri", do_scale=False,
                            reward_iteration_only=False, **KW)
    env.reset()
    lam0 = np.array(env.get_attr("lam"))
    assert_same(lam0, host_rng.lambda_stream(5, np.arange(n), 0, (-100, 0), (-10, 0)))
    act = rng.uniform(0, 0.6, (n, num_actions(M, "lower_tri")))
    obs, rew, done, infos = env.step(act)
    assert done.all()
    u, r = exact.reset(Q, 1.0, lam0)
    niter = np.zeros(n, np.int32)
    out = exact.step("sdc-v0", Q, 1.0, lam0, u, r, niter, r.copy(), act, prec_type="lower_tri", do_scale=False,
                     reward_strategy="residual_change")
    assert np.array_equal(infos.niter, niter)
    assert_same(infos.residual, out["resnorm"]); assert_reward_close(rew, out["reward"])
    assert_same(infos.lam, lam0)
    term = infos.terminal_observations()
    assert_same(term[:, 0], u); assert_same(term[:, 1], r)
    lam1 = host_rng.lambda_stream(5, np.arange(n), 1, (-100, 0), (-10, 0))
    assert_same(np.array(env.get_attr("lam")), lam1)
    u1, r1 = exact.reset(Q, 1.0, lam1)
    assert_same(obs[:, 0], u1); assert_same(obs[:, 1], r1)
    assert env.envs[0].num_episodes == 2 and env.envs[n - 1].niter == 0
    # a second step continues from the reset state (resnorm / niter / rng counters were rewritten consistently)
    act2 = rng.uniform(0, 0.6, act.shape)
    _, rew2, _, infos2 = env.step(act2)
    niter2 = np.zeros(n, np.int32)
    out2 = exact.step("sdc-v0", Q, 1.0, lam1, u1, r1, niter2, r1.copy(), act2, prec_type="lower_tri", do_scale=False,
                      reward_strategy="residual_change")
    assert np.array_equal(infos2.niter, niter2); assert_same(infos2.residual, out2["resnorm"])
    assert_reward_close(rew2, out2["reward"])
    assert_same(np.array(env.get_attr("lam")), host_rng.lambda_stream(5, np.arange(n), 2, (-100, 0), (-10, 0)))


def test_v1_autoreset_rollout_dense_m9():
    """sdc-v1 with per-env auto-reset (teams of one warp finish at different steps) against oracle envs"""
    n, M = 300, 9
    rng = np.random.default_rng(9)
    Q = collocation_matrix(M)
    Qd = fixed_preconditioner("LU", M, Q)
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, M=M, seed=21, prec="LU", reward_iteration_only=False,
                            norm_factor=3.0, **KW)
    obs = env.reset()
    draws = np.zeros(n, np.int64)
    lam = host_rng.lambda_stream(21, np.arange(n), draws, (-100, 0), (-10, 0))
    u, r = exact.reset(Q, 1.0, lam)
    rinit, niter = r.copy(), np.zeros(n, np.int32)
    assert_same(obs[:, 0], u); assert_same(obs[:, 1], r)
    ndone = 0
    for s in range(90):
        obs, rew, done, infos = env.step(np.zeros((n, M)))
        out = exact.step("sdc-v1", Q, 1.0, lam, u, r, niter, rinit, None, prec_type="fixed", Qd_fixed=Qd,
                         reward_strategy="residual_change", norm_factor=3.0)
        assert np.array_equal(done, out["done"]), f"step {s}"
        assert np.array_equal(infos.niter, niter); assert_same(infos.residual, out["resnorm"])
        assert_reward_close(rew, out["reward"])
        if done.any():
            term = infos.terminal_observations()
            assert_same(term[done, 0], u[done]); assert_same(term[done, 1], r[done])
            draws[done] += 1
            lam = np.where(done, host_rng.lambda_stream(21, np.arange(n), draws, (-100, 0), (-10, 0)), lam)
            nu, nr = exact.reset(Q, 1.0, lam)
            u[done], r[done], rinit[done], niter[done] = nu[done], nr[done], nr[done], 0
            ndone += int(done.sum())
        assert_same(obs[:, 0], u, f"step {s} obs u"); assert_same(obs[:, 1], r, f"step {s} obs r")
    assert ndone > n


@pytest.mark.parametrize("M", [8, 9])
@pytest.mark.parametrize("n", [1, 77])
def test_padding_is_untouched(M, n):
    """the padding [n, ld) of every plane doubles as a canary region: no lane of a partial team writes past N"""
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=1, prec_type="lower_tri", do_scale=False, **KW)
    env.reset()
    rng = np.random.default_rng(0)
    env.step(rng.uniform(0, 0.6, (n, num_actions(M, "lower_tri"))))
    for name in ("lam", "S", "resnorm", "niter", "episodes", "rng_ctr", "reward", "flags", "info_residual",
                 "info_niter", "terminal"):
        t = getattr(env, name)
        assert bool((t[..., n:] == 0).all()), f"{name}: write beyond the batch"
    assert bool((env.info_lam[n:] == 0).all())
