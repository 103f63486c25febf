"""GPU: ``sdc-v4`` (``SDCForceVecEnv``) against episodes of the reference's ``SDC_Full_Force_Env`` recorded in
``tests/golden/sdc_force_golden.npz`` (``tests/golden/make_golden_force.py``: the unmodified reference class with the
one repaired ``reward_func`` call), plus the DummyVecEnv auto-reset semantics on a larger batch."""
import os

import numpy as np
import pytest

import sdc_gym_b200
from sdc_gym_b200 import _lib
from tests.helpers import assert_reward_close, assert_same

pytestmark = pytest.mark.gpu
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sdc_force_golden.npz"))
CASES = {"m3_iter": dict(M=3), "m5_iter": dict(M=5), "m5_reschange": dict(M=5, reward_iteration_only=False),
         "m5_min": dict(M=5, prec="min"), "m7_fast": dict(M=7, reward_strategy="fast_convergence")}


@pytest.mark.parametrize("name", sorted(CASES))
def test_force_env_replays_reference_episodes(name):
    import torch
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    n, T, M = g["actions"].shape
    env = sdc_gym_b200.make("sdc-v4", num_envs=n, autoreset=False, output="torch", **KW, **CASES[name])
    obs = env.reset(lam=g["lam"]).cpu().numpy()
    assert_same(obs[:, 0], g["r0"], "initial residual")
    assert not obs[:, 1].any()
    checked = 0
    for t in range(T):
        v = g["valid"][:, t]
        if not v.any():
            break
        act = torch.as_tensor(g["actions"][:, t], device=env.device)
        out = env.step_tensor(act)
        o = out["obs"].cpu().numpy()
        assert_same(o[v, 0], g["res"][v, t], f"residual, try {t}")
        assert_same(o[v, 1], g["diag"][v, t], f"diagonal, try {t}")
        assert np.array_equal(out["niter"].cpu().numpy()[v], g["niter"][v, t])
        assert np.array_equal(out["ntries"].cpu().numpy()[v], g["ntries"][v, t])
        assert_same(out["residual"].cpu().numpy()[v], g["resnorm"][v, t], f"||r||, try {t}")
        assert np.array_equal(out["done"].cpu().numpy()[v], g["done"][v, t])
        assert_reward_close(out["reward"].cpu().numpy()[v], g["reward"][v, t])
        checked += int(v.sum())
    assert checked == int(g["valid"].sum()) and checked >= n


def test_force_env_autoreset_follows_dummy_vec_env():
    """done envs: terminal observation kept, new lambda drawn (one Philox draw, one episode counted), state
    (r0(new lambda), zeros), ntries 0; running envs: lambda, episode and draw counters untouched by the state restart"""
    import torch
    from sdc_gym_b200.precond import fixed_preconditioner
    n, M = 4096, 5
    env = sdc_gym_b200.make("sdc-v4", num_envs=n, M=M, seed=7, output="torch", **KW)
    env.reset()
    x = torch.as_tensor(np.diag(fixed_preconditioner("min", M)).copy(), device=env.device)
    gen = torch.Generator(device=env.device); gen.manual_seed(3)
    total_done = 0
    for s in range(6):
        lam0 = env.venv.lam[:, :n].clone()
        ep0, ctr0 = env.venv.episodes[:n].clone(), env.venv.rng_ctr[:n].clone()
        tries0 = env.ntries.clone()
        scaled = x[None] * (0.5 if s < 2 else 0.0) + torch.rand((n, M), dtype=torch.float64, device=env.device,
                                                                generator=gen) * 0.004
        out = env.step_tensor(2 * scaled - 1)
        done = out["done"]
        total_done += int(done.sum())
        assert torch.equal(out["ntries"], tries0 + 1)
        assert torch.equal(env.ntries, torch.where(done, torch.zeros_like(tries0), tries0 + 1))
        assert torch.equal(env.venv.episodes[:n], ep0 + done.to(ep0.dtype))
        assert torch.equal(env.venv.rng_ctr[:n], ctr0 + done.to(ctr0.dtype))
        assert torch.equal(env.venv.lam[:, :n][:, ~done], lam0[:, ~done])
        assert bool((out["lam"].real == lam0[0]).all()) and bool((out["lam"].imag == lam0[1]).all())
        obs, term = out["obs"], out["terminal"]
        assert torch.equal(obs[~done], term[~done])
        assert not bool(obs[done][:, 1].abs().any())
        # r0 of the new lambdas: a plain sdc-v0 env reset with them injected
        ref = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, output="torch", autoreset=False, **KW)
        lam_now = (env.venv.lam[0, :n] + 1j * env.venv.lam[1, :n]).cpu().numpy()
        r0 = ref.reset(lam=lam_now)[:, 1]
        assert torch.equal(obs[done][:, 0], r0[done])
        conv = out["converged"]
        assert torch.equal(done, conv | (tries0 + 1 >= 50))
    assert total_done > n // 2  # the weights reach the MIN diagonal on the second try


def test_force_env_host_api_and_registry():
    n, M = 64, 3
    env = sdc_gym_b200.make("sdc-v4", num_envs=n, M=M, seed=1, **KW)
    assert sdc_gym_b200.REGISTRY["sdc-v4"] == ("SDC_Full_Force_Env", 50)
    obs = env.reset()
    assert obs.shape == (n, 2, M) and obs.dtype == np.complex128 and env.observation_space.shape == (2, M)
    rng = np.random.default_rng(0)
    seen_done = False
    for _ in range(4):
        obs, rew, done, infos = env.step(rng.uniform(-1, -0.5, (n, M)))
        assert obs.shape == (n, 2, M) and rew.shape == (n,) and done.dtype == bool and len(infos) == n
        i0 = infos[0]
        assert set(i0) >= {"residual", "niter", "ntries", "lam"} and 1 <= i0["niter"] <= 50
        for i in np.flatnonzero(done)[:3]:
            seen_done = True
            assert infos[int(i)]["terminal_observation"].shape == (2, M)
            assert not obs[i, 1].any()
    assert seen_done
    with pytest.raises(NotImplementedError):
        sdc_gym_b200.make("sdc-v4", num_envs=2, M=3, collect_states=True, **KW)
