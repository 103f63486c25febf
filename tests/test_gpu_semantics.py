"""GPU: VecEnv semantics of the batched env - fused auto-reset (DummyVecEnv), the Philox lambda stream, lazy
infos / terminal observations, pipelined host step == device step, layout kernels, spectral radius, and
size-independent properties at the full benchmark size (2^20 envs)."""
import ctypes

import numpy as np
import pytest

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib, rng as host_rng
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner, num_actions, qdmat_from_output
from tests.helpers import assert_reward_close, assert_same

pytestmark = pytest.mark.gpu

KW = dict(M=5, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)


def good_actions(rng, n, M=5, spread=0.02):
    x = np.diag(fixed_preconditioner("min", M))
    return 2 * (x[None] + rng.uniform(-spread, spread, (n, M))) - 1


def test_reset_draws_the_philox_stream_and_is_shard_invariant():
    n = 3000
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=11, **KW)
    obs = env.reset()
    lam = np.array(env.get_attr("lam"))
    want = host_rng.lambda_stream(11, np.arange(n), 0, (-100, 0), (-10, 0))
    assert_same(lam, want, "lambda stream")
    u, r = exact.reset(collocation_matrix(5), 1.0, lam)
    assert_same(obs[:, 0], u); assert_same(obs[:, 1], r)
    # a shard starting at global env 1000 sees the same lambdas
    shard = sdc_gym_b200.make("sdc-v0", num_envs=500, seed=11, env_offset=1000, **KW)
    shard.reset()
    assert_same(np.array(shard.get_attr("lam")), want[1000:1500], "shard lambda stream")
    assert env.envs[0].num_episodes == 1 and env.envs[n - 1].niter == 0
    env.reset()
    assert_same(np.array(env.get_attr("lam")), host_rng.lambda_stream(11, np.arange(n), 1, (-100, 0), (-10, 0)))
    assert env.envs[5].num_episodes == 2


def test_v0_autoreset_matches_dummy_vec_env_protocol():
    n = 2048
    rng = np.random.default_rng(3)
    Q = collocation_matrix(5)
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=5, reward_iteration_only=False, **KW)
    obs0 = env.reset()
    lam0 = np.array(env.get_attr("lam"))
    act = good_actions(rng, n)
    obs, rew, done, infos = env.step(act)
    assert done.all() and obs.shape == (n, 2, 5) and obs.dtype == np.complex128
    # oracle solve on the first lambdas
    u, r = exact.reset(Q, 1.0, lam0)
    niter = np.zeros(n, np.int32)
    out = exact.step("sdc-v0", Q, 1.0, lam0, u, r, niter, r.copy(), act, reward_strategy="residual_change")
    assert np.array_equal(infos.niter, niter)
    assert_same(infos.residual, out["resnorm"]); assert_reward_close(rew, out["reward"])
    assert_same(infos.lam, lam0, "info['lam'] is the finished lambda")
    term = infos.terminal_observations()
    assert_same(term[:, 0], u); assert_same(term[:, 1], r)
    i7 = infos[7]
    assert i7["TimeLimit.truncated"] is False and i7["niter"] == niter[7]
    assert_same(i7["terminal_observation"], np.stack([u[7], r[7]]))
    # the returned observation is already the reset state of the NEXT lambda (draw index 1)
    lam1 = host_rng.lambda_stream(5, np.arange(n), 1, (-100, 0), (-10, 0))
    assert_same(np.array(env.get_attr("lam")), lam1)
    u1, r1 = exact.reset(Q, 1.0, lam1)
    assert_same(obs[:, 0], u1); assert_same(obs[:, 1], r1)
    assert env.envs[0].num_episodes == 2 and env.envs[0].niter == 0
    assert_same(env.envs[3].initial_residual, r1[3])
    assert_same(env.envs[3].state[1], r1[3])
    # success predicate of rl_playground.test_model
    succ = sum(1 for e_, i_ in zip(env.envs[:64], infos[:64]) if i_["niter"] < 50 and i_["residual"] < e_.restol)
    assert succ == int(((niter[:64] < 50) & (out["resnorm"][:64] < 1e-10)).sum())


def test_v1_autoreset_rollout_against_restated_dummy_vec_env():
    """200 steps of sdc-v1 with per-env auto-reset against a loop over oracle envs fed the same lambdas."""
    n, M = 512, 5
    rng = np.random.default_rng(9)
    Q = collocation_matrix(M)
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=21, reward_iteration_only=False, **KW)
    obs = env.reset()
    draws = np.zeros(n, np.int64)
    lam = host_rng.lambda_stream(21, np.arange(n), draws, (-100, 0), (-10, 0))
    u, r = exact.reset(Q, 1.0, lam)
    rinit, niter = r.copy(), np.zeros(n, np.int32)
    assert_same(obs[:, 0], u); assert_same(obs[:, 1], r)
    ndone = 0
    for s in range(120):
        act = good_actions(rng, n, spread=0.15)
        obs, rew, done, infos = env.step(act)
        out = exact.step("sdc-v1", Q, 1.0, lam, u, r, niter, rinit, act, reward_strategy="residual_change")
        assert np.array_equal(done, out["done"]), f"step {s}"
        assert np.array_equal(infos.niter, niter); assert_same(infos.residual, out["resnorm"])
        assert_reward_close(rew, out["reward"]); assert_same(infos.lam, lam)
        if done.any():
            term = infos.terminal_observations()
            assert_same(term[done, 0], u[done]); assert_same(term[done, 1], r[done])
            idx = np.nonzero(done)[0]
            assert "terminal_observation" in infos[int(idx[0])]
            trunc = [("TimeLimit.truncated" in infos[int(i)]) for i in idx]
            assert trunc == [bool(niter[i] >= 50) for i in idx]
            draws[done] += 1
            lam = np.where(done, host_rng.lambda_stream(21, np.arange(n), draws, (-100, 0), (-10, 0)), lam)
            nu, nr = exact.reset(Q, 1.0, lam)
            u[done], r[done], rinit[done], niter[done] = nu[done], nr[done], nr[done], 0
            ndone += int(done.sum())
        assert_same(obs[:, 0], u, f"step {s} obs u"); assert_same(obs[:, 1], r, f"step {s} obs r")
    assert ndone > n  # every env finished at least a couple of episodes


def test_pipelined_host_step_equals_device_step():
    n = 300_000  # several pipeline chunks
    rng = np.random.default_rng(1)
    act = rng.uniform(-1, 1, (n, 5))
    a = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=2, pipeline_chunks=5, **KW)
    b = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=2, **KW)
    c = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=2, pipeline_chunks=1, **KW)  # one transfer, caller's stream
    import torch
    oa = a.reset(); b.reset(); c.reset()
    for _ in range(2):
        obs, rew, done, infos = a.step(act)
        obs_c, rew_c, done_c, infos_c = c.step(act)
        assert_same(obs, obs_c); assert_same(rew, rew_c); assert np.array_equal(infos.niter, infos_c.niter)
        assert_same(infos.lam, infos_c.lam); assert np.array_equal(infos.flags, infos_c.flags)
        out = b.step_tensor(torch.as_tensor(act, device=b.device))
        assert_same(rew, out["reward"].cpu().numpy()); assert np.array_equal(infos.niter, out["niter"].cpu().numpy())
        assert_same(infos.residual, out["residual"].cpu().numpy())
        assert_same(obs, b.observation_tensor().cpu().numpy())
        assert_same(infos.lam, out["lam"].cpu().numpy())


def test_collect_states_autoreset_buffers():
    n = 64
    rng = np.random.default_rng(4)
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=8, collect_states=True, **KW)
    obs = env.reset()
    assert obs.shape == (n, 10, 50) and np.all(obs[:, :, 1:] == 0) and np.all(obs[:, :5, 0] == 1)
    for s in range(1, 30):
        obs, rew, done, infos = env.step(good_actions(rng, n))
        live = ~done
        assert np.all(obs[live][:, :, s + 1:] == 0) if s + 1 < 50 else True
        if done.any():
            t = infos.terminal_observations()
            assert t.shape == (n, 10, 50)
            assert np.all(obs[done][:, :, 1:] == 0)  # reset buffer
            break
    assert done.any()


def test_export_import_roundtrip_refresh_and_sum():
    import torch
    L = _lib.load()
    n, M = 1000, 7
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, **{**KW, "M": M})
    env.reset()
    rng = np.random.default_rng(0)
    u = rng.normal(size=(n, M)) + 1j * rng.normal(size=(n, M))
    r = rng.normal(size=(n, M)) + 1j * rng.normal(size=(n, M))
    env.set_state(u, r)
    snap = env._snapshot()
    assert_same(snap["obs"][:, 0], u); assert_same(snap["obs"][:, 1], r)
    assert_same(env.resnorm[:n].cpu().numpy(), np.abs(r).max(axis=1))
    x = torch.as_tensor(rng.normal(size=123457), device=env.device)
    out = torch.zeros(1, dtype=torch.float64, device=env.device)
    scratch = torch.zeros(L.sdcgym_sum_scratch_doubles(), dtype=torch.float64, device=env.device)
    _lib.check(L.sdcgym_sum_f64(x.numel(), x.data_ptr(), scratch.data_ptr(), out.data_ptr(), None), "sum")
    torch.cuda.synchronize()
    assert abs(out.item() - float(np.sum(x.cpu().numpy()))) < 1e-9


def test_err_paths_nan_and_divergence():
    n = 256
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=3, autoreset=False, **KW)
    env.reset()
    u = np.ones((n, 5), np.complex128)
    r = np.ones((n, 5), np.complex128)
    r[::2, 2] = np.nan
    env.set_state(u, r)
    _, rew, done, infos = env.step(np.zeros((n, 5)))
    assert done[::2].all() and np.all((infos.flags[::2] & 4) != 0) and np.all(rew[::2] == -0.1 * 51)
    assert np.all(np.isnan(infos.residual[::2]))
    # zero preconditioner on stiff lambdas diverges: err flag + penalty reward, like the reference
    env0 = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=3, prec="zeros", autoreset=False, **KW)
    env0.reset(lam=np.full(n, -90.0 - 5j))
    _, rew, done, infos = env0.step(None)
    assert np.all((infos.flags & 4) != 0) and np.all(rew == -0.1 * 51) and np.all(infos.niter < 50)


@pytest.mark.parametrize("prec_type", ["diag", "lower_diag", "lower_tri", "strictly_lower_tri"])
@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7, 8, 9])
def test_spectral_radius_against_lapack(M, prec_type):
    from sdc_gym_b200.loss import SpectralRadiusLoss
    rng = np.random.default_rng(M)
    Q = collocation_matrix(M)
    n, A = (2000 if M in (3, 5, 7) else 600), num_actions(M, prec_type)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    for cplx in (False, True):
        hi = 0.6 if prec_type in ("diag", "lower_diag") else 0.2
        qd = rng.uniform(0, hi, (n, A)) + (1j * rng.uniform(-0.1, 0.1, (n, A)) if cplx else 0)
        loss = SpectralRadiusLoss(M, 1.0, prec_type)
        rho = loss.spectral_radii(lam.reshape(-1, 1), qd).cpu().numpy()
        ref, nk = np.empty(n), np.empty(n)
        for i in range(n):
            Qd = qdmat_from_output(qd[i], M, prec_type)
            K = lam[i] * np.linalg.inv(np.eye(M) - lam[i] * Qd) @ (Q - Qd)
            ref[i] = max(abs(np.linalg.eigvals(K)))
            nk[i] = np.linalg.norm(K, 2)
        # 1e-10 relative.  Strongly non-normal K (large M, strictly lower triangular Q_delta: ||K||_2 up
        # to 1e8 with rho 1e7, cond(P) 1e8) - there LAPACK's own inv + eigvals are only good to ~1e-9 ||K||_2, and so
        # is the comparison
        tol = 1e-10 * ref + 1e-13 * ref.max() + np.where(nk > 10.0 * ref, 1e-9 * nk, 0.0)
        assert np.all(np.abs(rho - ref) <= tol), np.max(np.abs(rho - ref) / ref)
        assert abs(float(loss(lam, qd)) - ref.mean()) <= 1e-10 * ref.mean() + tol.mean()


def test_spectral_radius_grid_and_fixed_prec():
    from sdc_gym_b200.loss import SpectralRadiusLoss
    M = 5
    Q = collocation_matrix(M)
    x = np.diag(fixed_preconditioner("min", M))
    loss = SpectralRadiusLoss(M, 1.0, "diag")
    g = loss.grid(33, 17, [-100, 0], [-10, 0], x).cpu().numpy()
    re, im = np.linspace(-100, 0, 33), np.linspace(-10, 0, 17)
    for a in (0, 7, 32):
        for b in (0, 5, 16):
            l = complex(re[a], im[b])
            ref = max(abs(np.linalg.eigvals(l * np.linalg.inv(np.eye(M) - l * np.diag(x)) @ (Q - np.diag(x)))))
            assert abs(g[a, b] - ref) <= 1e-10 * max(ref, 1e-3)
    # row shards of the grid (one per rank in a multi-GPU run) are the rows of the full grid, bit for bit
    parts = [loss.grid(33, 17, [-100, 0], [-10, 0], x, rows=r).cpu().numpy() for r in ((0, 11), (11, 12), (12, 12), (12, 33))]
    assert [p.shape for p in parts] == [(11, 17), (1, 17), (0, 17), (21, 17)]
    assert np.array_equal(np.concatenate(parts), g)
    assert float(loss.grid_mean(33, 17, [-100, 0], [-10, 0], x)) == pytest.approx(g.mean(), rel=1e-14)
    with pytest.raises(ValueError):
        loss.grid(33, 17, [-100, 0], [-10, 0], x, rows=(5, 34))
    for prec in ("LU", "min", "EE", "zeros"):
        lossf = SpectralRadiusLoss(M, 1.0, prec=prec)
        lam = np.array([-50 - 3j, -1 - 0.5j, -99.5 - 9j])
        Qd = fixed_preconditioner(prec, M, Q)
        ref = np.array([max(abs(np.linalg.eigvals(l * np.linalg.inv(np.eye(M) - l * Qd) @ (Q - Qd)))) for l in lam])
        assert np.allclose(lossf.spectral_radii(lam).cpu().numpy(), ref, rtol=1e-10, atol=1e-12)


def test_full_size_properties_one_million_envs():
    """BASELINE config[1] size: properties that need no oracle run, plus an oracle check on a 8192-env subsample."""
    import torch
    n = 1 << 20
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=0, **KW)
    env.reset()
    lam0 = env.lam[:, :n].clone()
    gen = torch.Generator(device=env.device); gen.manual_seed(1)
    act = torch.rand((n, 5), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1
    act[: n // 2] = torch.as_tensor(good_actions(np.random.default_rng(0), n // 2, spread=0.05), device=env.device)
    out = env.step_tensor(act)
    niter, res, flags, rew = (out[k].clone() for k in ("niter", "residual", "flags", "reward"))
    conv, err = (flags & 2) != 0, (flags & 4) != 0
    assert bool(((flags & 1) != 0).all())
    assert bool(((niter >= 1) & (niter <= 50)).all())
    assert bool((res[conv] < 1e-10).all()) and bool((res[~conv & ~err] >= 1e-10).all())
    assert bool((niter[~conv & ~err] == 50).all()) and not bool((conv & err).any())
    assert bool(torch.equal(rew[err], torch.full_like(rew[err], -0.1 * 51)))
    assert bool(torch.equal(rew[~err], niter[~err].double() * -0.1))
    assert 0.05 < float(conv.double().mean()) < 0.6  # part of the "good" half converges
    # determinism + sub-range launches: re-running any slice of the batch reproduces the same bits
    env2 = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=0, pipeline_chunks=7, **KW)
    env2.reset()
    assert bool(torch.equal(env2.lam[:, :n], lam0))
    o2, r2, d2, i2 = env2.step(act.cpu().numpy())
    assert np.array_equal(i2.niter, niter.cpu().numpy()) and assert_same(i2.residual, res.cpu().numpy()) is None
    # oracle on a strided subsample
    idx = np.arange(0, n, n // 8192)
    lam_h = (lam0[0] + 1j * lam0[1]).cpu().numpy()[idx]
    Q = collocation_matrix(5)
    u, r = exact.reset(Q, 1.0, lam_h)
    nit = np.zeros(len(idx), np.int32)
    o = exact.step("sdc-v0", Q, 1.0, lam_h, u, r, nit, r.copy(), act.cpu().numpy()[idx])
    assert np.array_equal(niter.cpu().numpy()[idx], nit)
    assert_same(res.cpu().numpy()[idx], o["resnorm"])
    term = out["terminal"].cpu().numpy()[:, idx]
    assert_same((term[0:10:2] + 1j * term[1:10:2]).T, u); assert_same((term[10::2] + 1j * term[11::2]).T, r)


@pytest.mark.parametrize("n", [0, 1, 31, 33, 129, 1000])
@pytest.mark.parametrize("M", [2, 9])
def test_ragged_batch_sizes_and_extreme_M(n, M):
    """empty, single-env and non-multiple-of-warp batches at the smallest / largest supported M"""
    rng = np.random.default_rng(n + M)
    Q = collocation_matrix(M)
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=1, autoreset=False, **{**KW, "M": M})
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    obs = env.reset(lam=lam)
    assert obs.shape == (n, 2, M)
    act = rng.uniform(-0.6, -0.2, (n, M))
    obs, rew, done, infos = env.step(act)
    assert obs.shape == (n, 2, M) and rew.shape == (n,) and done.shape == (n,) and len(infos) == n
    if n:
        u, r = exact.reset(Q, 1.0, lam)
        nit = np.zeros(n, np.int32)
        o = exact.step("sdc-v0", Q, 1.0, lam, u, r, nit, r.copy(), act)
        snap = env._snapshot()
        assert_same(snap["obs"][:, 0], u); assert_same(snap["obs"][:, 1], r)
        assert np.array_equal(infos.niter, nit) and done.all()
    # compute-sanitizer is closed on this pool: the padding [n, ld) of every plane doubles as a canary region
    for name in ("lam", "S", "resnorm", "niter", "episodes", "rng_ctr", "reward", "flags", "info_residual",
                 "info_niter", "terminal"):
        t = getattr(env, name)
        assert bool((t[..., n:] == 0).all()), f"{name}: write beyond the batch"
    assert bool((env.info_lam[n:] == 0).all())


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        sdc_gym_b200.make("sdc-v0", num_envs=4, **{**KW, "M": 10})
    with pytest.raises(NotImplementedError):  # float32 Q_delta + eigenvalue reward: not composed (vec_env.py)
        sdc_gym_b200.make("sdc-v0", num_envs=4, use_doubles=False, reward_strategy="spectral_radius", **KW)
    with pytest.raises(NotImplementedError):
        sdc_gym_b200.make("sdc-v0", num_envs=4, reward_strategy="nope", **KW)
    env = sdc_gym_b200.make("sdc-v0", num_envs=4, **KW)
    env.reset()
    with pytest.raises(ValueError):
        env.step(np.zeros((4, 3)))
    with pytest.raises(TypeError):
        env.step(np.zeros((4, 5), np.complex128))


def test_use_doubles_false_action_space_and_parity():
    """use_doubles=False (SAC, utils/utils.py:279-280): float32 / complex64 action space (sdc_env.py:95-110); the
    reference then forms lam*dt*Qdmat in complex64 (golden cases *_f32 / *_c64 pin the arithmetic); here a larger
    batch against the rounding-exact oracle, fed with float32 arrays like SB3 would."""
    n, M = 3000, 5
    rng = np.random.default_rng(21)
    Q = collocation_matrix(M)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    for kind in ("sdc-v0", "sdc-v1"):
        env = sdc_gym_b200.make(kind, num_envs=n, use_doubles=False, autoreset=False, **KW)
        assert env.action_space.dtype == np.float32 and env.action_space.shape == (M,)
        env.reset(lam=lam)
        u, r = exact.reset(Q, 1.0, lam)
        xmin = np.diag(fixed_preconditioner("min", M))
        act = (2 * (xmin[None] + rng.uniform(-0.02, 0.02, (n, M))) - 1).astype(np.float32)
        _, rew, done, infos = env.step(act)
        nit = np.zeros(n, np.int32)
        o = exact.step(kind, Q, 1.0, lam, u, r, nit, r.copy(), act.astype(np.float64), use_doubles=False)
        snap = env._snapshot()
        assert_same(snap["obs"][:, 0], u); assert_same(snap["obs"][:, 1], r)
        assert np.array_equal(infos.niter, nit)
        assert_same(infos.residual, o["resnorm"])
        # and it is a different result from the float64 action space on the same numbers
        u2, r2 = exact.reset(Q, 1.0, lam)
        exact.step(kind, Q, 1.0, lam, u2, r2, np.zeros(n, np.int32), r2.copy(), act.astype(np.float64))
        assert not np.array_equal(u, u2)
    envc = sdc_gym_b200.make("sdc-v1", num_envs=4, use_doubles=False, free_action_space=True, do_scale=False, **KW)
    assert envc.action_space.dtype == np.complex64


def test_zero_residual_reward_is_nan_where_the_reference_raises():
    """lambda = 0: C = I, r0 = 0 exactly; `math.log(0)` raises ValueError in the reference's residual_change reward
    (sdc_env.py:337-350) - documented deviation: NaN for that env; fast_convergence returns 1000 (:380)."""
    n = 64
    lam = np.zeros(n, np.complex128)
    lam[1::2] = -1.0 - 0.5j
    act = np.zeros((n, 5))
    for strategy, want in (("residual_change", np.nan), ("fast_convergence", 1000.0)):
        env = sdc_gym_b200.make("sdc-v1", num_envs=n, reward_iteration_only=None, reward_strategy=strategy,
                                autoreset=False, **KW)
        obs = env.reset(lam=lam)
        assert np.all(obs[0::2, 1] == 0)
        _, rew, done, infos = env.step(act)
        assert np.all(infos.residual[0::2] == 0.0) and np.all(np.isfinite(rew[1::2]))
        if np.isnan(want):
            assert np.all(np.isnan(rew[0::2]))
        else:
            # converged at step 1: bonus (51 - 1)^2 * 10 on top of the zero-norm reward 1000 (sdc_env.py:370-385)
            assert np.all(rew[0::2] == want * (50.0 ** 2) * 10.0)


def test_spectral_radius_reward_strategy():
    n, M = 500, 5
    rng = np.random.default_rng(12)
    Q = collocation_matrix(M)
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=2, reward_strategy="spectral_radius", autoreset=False, **KW)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    env.reset(lam=lam)
    act = good_actions(rng, n, spread=0.05)
    _, rew, done, infos = env.step(act)
    d = np.clip(0.5 * (act + 1), 0, 1)
    for i in range(0, n, 7):
        Qd = np.diag(d[i])
        ref = max(abs(np.linalg.eigvals(lam[i] * np.linalg.inv(np.eye(M) - lam[i] * Qd).dot(Q - Qd))))
        if infos.flags[i] & 4:
            assert rew[i] == -0.1 * 51
        else:
            assert abs(rew[i] - ref) <= 1e-10 * ref


def test_curriculum_lambda_interval():
    """lambda_real_interpolation_interval (sdc_env.py:287-292): the lower real bound follows np.interp of the
    episode count."""
    n = 2000
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=9, lambda_real_interpolation_interval=[2, 6], **KW)
    for ep in range(1, 9):
        env.reset()
        lam = np.array(env.get_attr("lam"))
        lo = float(np.interp(ep, [2, 6], [0, -100]))
        want = host_rng.lambda_stream(9, np.arange(n), ep - 1, (-100, 0), (-10, 0), re_lo_override=lo)
        assert_same(lam, want, f"episode {ep}")
        assert np.all(lam.real >= lo - 1e-12) and np.all(lam.real <= 0)
    env.set_num_episodes(100)
    env.reset()
    assert np.array(env.get_attr("lam")).real.min() < -90


def test_torch_output_mode_and_seed_contract():
    import torch
    n = 257
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=4, output="torch", **KW)
    ref = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=4, **KW)
    o_t, o_n = env.reset(), ref.reset()
    assert o_t.is_cuda and o_t.dtype == torch.complex128 and o_t.shape == (n, 2, 5)
    assert_same(o_t.cpu().numpy(), o_n)
    act = np.random.default_rng(0).uniform(-1, 1, (n, 5))
    for a in (act, torch.as_tensor(act), torch.as_tensor(act, device="cuda")):  # numpy, CPU tensor, CUDA tensor
        obs, rew, done, info = env.step(a)
        obs_n, rew_n, done_n, info_n = ref.step(act)
        assert obs.is_cuda and rew.is_cuda and done.dtype == torch.bool
        assert_same(obs.cpu().numpy(), obs_n); assert_same(rew.cpu().numpy(), rew_n)
        assert np.array_equal(done.cpu().numpy(), done_n) and np.array_equal(info["niter"].cpu().numpy(), info_n.niter)
    seeds = env.seed(10)
    assert len(seeds) == n and seeds[0] == 10 and seeds[n - 1] == 10 + n - 1
    assert env.seed(None)[0] is None
    env.step_async(act)
    assert env.step_wait()[0].shape == (n, 2, 5)
    assert len(env.get_attr("M")) == n and env.get_attr("restol", indices=[0, 5]) == [1e-10, 1e-10]
    env.env_method("set_num_episodes", 7)
    assert env.envs[3].num_episodes == 7
    env.set_attr("num_episodes", 9, indices=[3])
    assert env.envs[3].num_episodes == 9 and env.envs[4].num_episodes == 7
    env.close()
