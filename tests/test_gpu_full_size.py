"""GPU: the BASELINE.json configurations at their full per-GPU sizes, checked through size-independent properties
plus the oracle on a strided subsample (the oracle cannot run millions of envs in seconds)."""
import numpy as np
import pytest

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner, num_actions, qdmat_from_output
from tests.helpers import assert_same

pytestmark = pytest.mark.gpu
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)


@pytest.mark.parametrize("M", [3, 5, 7, 9])
@pytest.mark.parametrize("prec_type", ["lower_tri", "strictly_lower_tri"])
@pytest.mark.parametrize("phased", [False, True])
def test_config2_m_sweep_triangular_4m_envs(M, prec_type, phased):
    """config 2: M sweep 3/5/7/9 with lower_tri and strictly_lower_tri Q_delta, 4M envs on one GPU; single launch and
    phased solve (M <= 7)"""
    import torch
    if phased and M > 7:
        pytest.skip("M >= 8 runs the lane-team kernel")
    n = 1 << 22
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, prec_type=prec_type, do_scale=False, seed=M, phased=phased, **KW)
    assert env.phased == phased
    env.reset()
    lam0 = torch.view_as_complex(torch.stack([env.lam[0, :n], env.lam[1, :n]], dim=1).contiguous()).clone()
    A = num_actions(M, prec_type)
    gen = torch.Generator(device=env.device); gen.manual_seed(M)
    act = torch.rand((n, A), dtype=torch.float64, device=env.device, generator=gen) * 0.12
    out = env.step_tensor(act)
    niter, res, flags, rew = out["niter"], out["residual"], out["flags"], out["reward"]
    conv, err = (flags & 2) != 0, (flags & 4) != 0
    assert bool(((flags & 1) != 0).all()) and bool(((niter >= 1) & (niter <= 50)).all())
    assert bool((res[conv] < 1e-10).all()) and not bool((conv & err).any())
    assert bool((niter[~conv & ~err] == 50).all())
    assert bool(torch.equal(rew[~err], niter[~err].double() * -0.1))
    assert bool(torch.equal(out["lam"], lam0))
    idx = np.arange(0, n, n // 1024)
    lam_h = lam0.cpu().numpy()[idx]
    Q = collocation_matrix(M)
    u, r = exact.reset(Q, 1.0, lam_h)
    nit = np.zeros(len(idx), np.int32)
    o = exact.step("sdc-v0", Q, 1.0, lam_h, u, r, nit, r.copy(), act.cpu().numpy()[idx], prec_type=prec_type,
                   do_scale=False)
    assert np.array_equal(niter.cpu().numpy()[idx], nit)
    assert_same(res.cpu().numpy()[idx], o["resnorm"])
    term = out["terminal"][:, torch.as_tensor(idx, device=env.device)].cpu().numpy()
    assert_same((term[0:2 * M:2] + 1j * term[1:2 * M:2]).T, u)
    assert_same((term[2 * M::2] + 1j * term[2 * M + 1::2]).T, r)


def test_config3_spectral_radius_grid_4096():
    """config 3: rho over a 4096 x 4096 lambda grid, M=5 diag Q_delta"""
    import torch
    from sdc_gym_b200.loss import SpectralRadiusLoss
    M, G = 5, 4096
    Q = collocation_matrix(M)
    x = np.diag(fixed_preconditioner("min", M))
    loss = SpectralRadiusLoss(M, 1.0, "diag")
    rho = loss.grid(G, G, [-100, 0], [-10, 0], x)
    assert rho.shape == (G, G) and bool(torch.isfinite(rho).all()) and float(rho.min()) >= 0.0
    assert float(rho[-1, -1]) < 1e-12  # lambda = 0: K = 0
    mean = float(loss.mean(rho.reshape(-1)))
    assert abs(mean - float(rho.mean())) <= 1e-12 * mean
    re, im = np.linspace(-100, 0, G), np.linspace(-10, 0, G)
    rng = np.random.default_rng(0)
    rho_h = rho.cpu().numpy()
    for a, b in zip(rng.integers(0, G, 200), rng.integers(0, G, 200)):
        l = complex(re[a], im[b])
        ref = max(abs(np.linalg.eigvals(l * np.linalg.inv(np.eye(M) - l * np.diag(x)) @ (Q - np.diag(x)))))
        assert abs(rho_h[a, b] - ref) <= 1e-10 * max(ref, 1e-6)
    # the MIN preconditioner is a contraction on the whole box
    assert float(rho.max()) < 1.0
    # a learned complex diagonal: batch API on 1M samples against a subsample
    B = 1 << 20
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    lam = torch.complex(torch.rand(B, dtype=torch.float64, device="cuda", generator=gen) * -100,
                        torch.rand(B, dtype=torch.float64, device="cuda", generator=gen) * -10)
    outp = torch.complex(torch.rand((B, M), dtype=torch.float64, device="cuda", generator=gen) * 0.4,
                         (torch.rand((B, M), dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 0.1)
    r2 = loss.spectral_radii(lam.reshape(-1, 1), outp).cpu().numpy()
    lam_h, out_h = lam.cpu().numpy(), outp.cpu().numpy()
    for i in range(0, B, B // 100):
        Qd = qdmat_from_output(out_h[i], M, "diag")
        ref = max(abs(np.linalg.eigvals(lam_h[i] * np.linalg.inv(np.eye(M) - lam_h[i] * Qd) @ (Q - Qd))))
        assert abs(r2[i] - ref) <= 1e-10 * ref


def test_config4_rollout_collection_64m_env_steps():
    """config 4: sdc-v1 rollouts with device VecNormalize(norm_obs), 4M envs x 16 steps = 64M env-steps"""
    import torch
    from sdc_gym_b200.dist import RolloutStats
    n, T, M = 1 << 22, 16, 5
    venv = sdc_gym_b200.make("sdc-v1", num_envs=n, M=M, seed=3, reward_iteration_only=False, **KW)
    env = sdc_gym_b200.VecNormalize(venv, norm_obs=True, norm_reward=True)
    env.reset()
    x = torch.as_tensor(np.diag(fixed_preconditioner("min", M)), device=venv.device)
    gen = torch.Generator(device=venv.device); gen.manual_seed(5)
    stats = RolloutStats(venv.device)
    episodes0 = venv.episodes[:n].clone()
    for t in range(T):
        a = 2 * (x[None] + (torch.rand((n, M), dtype=torch.float64, device=venv.device, generator=gen) - 0.5) * 0.3) - 1
        out = env.step_tensor(a)
        stats.update(out)
        assert bool(((out["niter"] >= 1) & (out["niter"] <= 50)).all())
        assert bool((out["obs_planes"].abs() <= 10.0).all()) and bool((out["reward"].abs() <= 10.0).all())
    red = stats.reduce()
    assert red["env_steps"] == float(n * T) == float(1 << 26)
    assert red["episodes"] == float((venv.episodes[:n] - episodes0).sum().item())  # every done env was reset once
    assert red["converged"] + red["diverged"] <= red["episodes"] and red["episodes"] > 0
    assert abs(env.obs_rms.count - (n * (T + 1) + 1e-4)) < 1.0
    assert bool(torch.isfinite(env.obs_rms.mean).all()) and bool((env.obs_rms.var > 0).all())
    # unfinished envs have niter == number of steps since their last reset
    assert bool((venv.niter[:n] <= T).all())


def _oracle_chunk(args):
    lam, act = args
    Q = collocation_matrix(5)
    u, r = exact.reset(Q, 1.0, lam)
    nit = np.zeros(lam.shape[0], np.int32)
    o = exact.step("sdc-v0", Q, 1.0, lam, u, r, nit, r.copy(), act)
    return u, r, nit, o["resnorm"], o["done"], o["err"]


@pytest.mark.parametrize("sweep_mode", ["exact", "certified"])
def test_config1_every_env_of_the_headline_batch_against_the_oracle(sweep_mode):
    """config 1 at FULL size, every env (not a subsample): 2^20 sdc-v0 envs, M = 5, diagonal Q_delta, half of the batch
    with uniform random actions (the benchmark workload), half near the MIN preconditioner (about half of those
    converge).  The rounding-exact C oracle runs on all host cores (~10-30 s).  exact mode: every output bit-equal.
    certified mode: niter / converged / err bit-equal, states within 1e-12 of ||u0|| + ||C|| ||u||."""
    import multiprocessing as mp
    import os

    import torch
    n, M = 1 << 20, 5
    rng = np.random.default_rng(2027)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    act = rng.uniform(-1, 1, (n, M))
    x = np.diag(fixed_preconditioner("min", M))
    act[n // 2:] = 2 * (x[None] + rng.uniform(-0.03, 0.03, (n - n // 2, M))) - 1
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, autoreset=False, sweep_mode=sweep_mode, **KW)
    env.reset(lam=lam)
    out = env.step_tensor(torch.as_tensor(act, device=env.device))
    snap = env._snapshot()
    cores = os.cpu_count() or 1
    bounds = np.linspace(0, n, 4 * cores + 1).astype(int)
    with mp.get_context("spawn").Pool(cores) as pool:
        parts = pool.map(_oracle_chunk, [(lam[a:b], act[a:b]) for a, b in zip(bounds, bounds[1:])])
    u, r, nit, res, conv, err = (np.concatenate([p[k] for p in parts]) for k in range(6))
    flags = out["flags"].cpu().numpy()
    assert np.array_equal(out["niter"].cpu().numpy(), nit)
    assert np.array_equal((flags & 2) != 0, conv) and np.array_equal((flags & 4) != 0, err)
    assert 0.1 < conv.mean() < 0.3 and 0.05 < err.mean() < 0.2  # the workload exercises all three exits
    gu, gr, gres = snap["obs"][:, 0], snap["obs"][:, 1], out["residual"].cpu().numpy()
    if sweep_mode == "exact":
        assert_same(gu, u)
        assert_same(gr, r)
        assert_same(gres, res)
    else:
        scale = 1.0 + 100.0 * np.abs(u).max(axis=1)
        assert np.all(np.abs(gu - u).max(axis=1) <= 1e-12 * scale)
        assert np.all(np.abs(gr - r).max(axis=1) <= 1e-12 * scale)
        assert np.all(np.abs(gres - res) <= 1e-12 * scale)
        nfb, _ = env.fallback_stats()
        assert 0 < nfb < 0.1 * n
        fb = env.fallback_list[:nfb].cpu().numpy()
        assert_same(gu[fb], u[fb])  # re-run by the exact kernel: bit for bit
        assert_same(gr[fb], r[fb])
