"""GPU tests of the host-facing step (`SDCVecEnv.step(numpy actions)` -> `sdcgym_pipe_step_block`): result blocks,
ownership of the returned arrays, the constant-u elision of `sdc-v0`, every chunking regime, stale `infos`, seeding,
device placement.  The checker is the device-resident step (itself bit-compared with the oracle in
test_gpu_parity / test_gpu_semantics) and, where cheap, the oracle directly."""
import numpy as np
import pytest

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from tests.helpers import assert_reward_close, assert_same

pytestmark = pytest.mark.gpu

KW = dict(M=5, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)


@pytest.mark.parametrize("kind", ["sdc-v0", "sdc-v1"])
@pytest.mark.parametrize("n", [1, 8, 2000, 20_000, 50_000, 150_000, 500_000])
def test_host_step_equals_device_step_in_every_chunk_regime(kind, n):
    """1 transfer (< 32 k envs), 2, 4 and 8 pipeline chunks: same bits as the device-resident step"""
    import torch

    rng = np.random.default_rng(n)
    a = sdc_gym_b200.make(kind, num_envs=n, seed=5, **KW)
    b = sdc_gym_b200.make(kind, num_envs=n, seed=5, **KW)
    a.reset(); b.reset()
    for s in range(3):
        act = rng.uniform(-1, 1, (n, 5))
        obs, rew, done, infos = a.step(act)
        out = b.step_tensor(torch.as_tensor(act, device=b.device))
        assert obs.shape == (n, 2, 5) and obs.dtype == np.complex128
        assert_same(obs, b.observation_tensor().cpu().numpy(), f"{kind} n={n} step {s} obs")
        assert_same(rew, out["reward"].cpu().numpy()); assert np.array_equal(infos.niter, out["niter"].cpu().numpy())
        assert_same(infos.residual, out["residual"].cpu().numpy()); assert_same(infos.lam, out["lam"].cpu().numpy())
        assert np.array_equal(infos.flags, out["flags"].cpu().numpy())
        assert np.array_equal(done, (out["flags"] & 1).bool().cpu().numpy())
        if done.any():
            term = infos.terminal_observations()
            tb = out["terminal"].cpu().numpy()
            tb = np.stack([(tb[0:10:2] + 1j * tb[1:10:2]).T, (tb[10::2] + 1j * tb[11::2]).T], axis=1)
            assert_same(term[done], tb[done])
    if kind == "sdc-v0":
        assert np.all(obs[:, 0] == 1.0)  # the reset state of the next lambda: u = 1 (never transferred)


def test_returned_arrays_stay_valid_while_the_caller_holds_them():
    """DummyVecEnv hands out arrays the caller owns.  Here a step writes into a result block only when nothing outside
    the env references its arrays any more; a caller that keeps results alive gets further blocks, then copies."""
    n = 4096
    rng = np.random.default_rng(0)
    env = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, max_host_sets=3, **KW)
    env.reset()
    kept = []
    for s in range(6):
        obs, rew, done, infos = env.step(rng.uniform(-1, 1, (n, 5)))
        kept.append((obs, rew, infos, obs.copy(), rew.copy(), infos.niter.copy(), infos.lam.copy()))
        for o, r, i, oc, rc, nc, lc in kept:  # nothing handed out earlier has been overwritten
            assert np.array_equal(o, oc) and np.array_equal(r, rc) and np.array_equal(i.niter, nc) and np.array_equal(i.lam, lc)
    assert len(env._host["sets"]) == 3 and env.host_set_copies == 3  # steps 4..6 had to copy
    # the usual loop (results rebound every step) ping-pongs between two blocks and never copies
    env2 = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, **KW)
    env2.reset()
    rng = np.random.default_rng(0)
    for s in range(6):
        obs, rew, done, infos = env2.step(rng.uniform(-1, 1, (n, 5)))
        assert np.array_equal(obs, kept[s][3]) and np.array_equal(rew, kept[s][4])
    assert len(env2._host["sets"]) == 2 and env2.host_set_copies == 0
    # a view derived from a result pins its block as well
    v = obs[7, 1]
    del obs, rew, done, infos
    keep = v.copy()
    for s in range(3):
        env2.step(rng.uniform(-1, 1, (n, 5)))
    assert np.array_equal(v, keep)
    # reuse_buffers=True: one block, overwritten in place (the caller opted out of ownership)
    env3 = sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, reuse_buffers=True, **KW)
    env3.reset()
    o1 = env3.step(rng.uniform(-1, 1, (n, 5)))[0]
    o2 = env3.step(rng.uniform(-1, 1, (n, 5)))[0]
    assert o1 is o2 and len(env3._host["sets"]) == 1


def test_v0_host_observation_is_reset_state_against_oracle():
    n = 3000
    rng = np.random.default_rng(3)
    Q = collocation_matrix(5)
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=9, **KW)
    obs0 = env.reset()
    lam = env._snapshot()["lam"].copy()
    act = rng.uniform(-1, 1, (n, 5))
    obs, rew, done, infos = env.step(act)
    u, r = exact.reset(Q, 1.0, lam)
    assert_same(obs0[:, 0], u); assert_same(obs0[:, 1], r)
    nit = np.zeros(n, np.int32)
    out = exact.step("sdc-v0", Q, 1.0, lam, u, r, nit, r.copy(), act)
    assert done.all() and np.array_equal(infos.niter, nit) and np.array_equal(infos.lam, lam)
    assert_same(infos.residual, out["resnorm"]); assert_reward_close(rew, out["reward"])
    term = infos.terminal_observations()
    assert_same(term[:, 0], u); assert_same(term[:, 1], r)
    # the returned observation is the reset state of the NEXT lambda (DummyVecEnv auto-reset)
    lam2 = env._snapshot()["lam"]
    u2, r2 = exact.reset(Q, 1.0, lam2)
    assert_same(obs[:, 0], u2); assert_same(obs[:, 1], r2)
    assert not obs.flags.c_contiguous and np.ascontiguousarray(obs).shape == (n, 2, 5)
    import torch

    assert torch.from_numpy(obs).shape == (n, 2, 5)  # positive strides: usable as a tensor without a copy


def test_stale_infos_raise_instead_of_returning_a_later_steps_terminal_observation():
    n = 64
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=2, **KW)
    env.reset()
    _, _, _, infos = env.step(np.zeros((n, 5)))
    first = infos[0]["terminal_observation"].copy()  # fetched in time: cached on the infos object
    env.step(np.zeros((n, 5)))
    assert np.array_equal(infos[0]["terminal_observation"], first)
    _, _, _, infos2 = env.step(np.zeros((n, 5)))
    env.step(np.zeros((n, 5)))
    with pytest.raises(RuntimeError, match="stepped since"):
        infos2[0]
    # keep_terminal=False: no terminal planes are written at all
    env3 = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=2, keep_terminal=False, **KW)
    env3.reset()
    o, r, d, i3 = env3.step(np.zeros((n, 5)))
    with pytest.raises(RuntimeError, match="keep_terminal"):
        i3.terminal_observations()


def test_seed_none_draws_fresh_entropy_and_seed_used_reproduces():
    n = 256
    a = sdc_gym_b200.make("sdc-v0", num_envs=n, **KW)
    b = sdc_gym_b200.make("sdc-v0", num_envs=n, **KW)
    oa, ob = a.reset(), b.reset()
    assert a.seed_used != b.seed_used and not np.array_equal(oa, ob)
    c = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=a.seed_used, **KW)
    assert np.array_equal(c.reset(), oa)
    a.seed(None)
    assert a.seed_used != c.seed_used


def test_env_on_a_device_that_is_not_current():
    """the C ABI launches on the thread's current device: every public method must switch to the env's device
    (allocations honoured `device=`, launches did not); also the > 48 KB shared-memory opt-in per device"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 5000
    rng = np.random.default_rng(0)
    act = rng.uniform(-1, 1, (n, 5))
    torch.cuda.set_device(0)
    e0 = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=4, device="cuda:0", **KW)
    e1 = sdc_gym_b200.make("sdc-v0", num_envs=n, seed=4, device="cuda:1", **KW)  # built and driven while cuda:0 is current
    o0, o1 = e0.reset(), e1.reset()
    assert np.array_equal(o0, o1) and torch.cuda.current_device() == 0
    r0, r1 = e0.step(act), e1.step(act)
    assert np.array_equal(r0[0], r1[0]) and np.array_equal(r0[1], r1[1]) and np.array_equal(r0[3].niter, r1[3].niter)
    t1 = e1.step_tensor(torch.as_tensor(act, device="cuda:1"))
    t0 = e0.step_tensor(torch.as_tensor(act, device="cuda:0"))
    assert torch.equal(t0["niter"].cpu(), t1["niter"].cpu()) and t1["niter"].device.index == 1
    assert torch.cuda.current_device() == 0
    from sdc_gym_b200.loss import SpectralRadiusLoss

    l1 = SpectralRadiusLoss(5, 1.0, "diag", device="cuda:1")
    l0 = SpectralRadiusLoss(5, 1.0, "diag", device="cuda:0")
    lam = rng.uniform(-100, 0, 64) + 1j * rng.uniform(-10, 0, 64)
    d = rng.uniform(0, 1, (64, 5))
    assert torch.equal(l0.spectral_radii(lam, d).cpu(), l1.spectral_radii(lam, d).cpu())
    # the phased dense solve (two more kernels with a > 48 KB shared-memory opt-in) on the device that is not current
    from sdc_gym_b200.precond import num_actions

    n2 = 20000
    kw2 = dict(M=5, prec_type="strictly_lower_tri", do_scale=False, seed=2, **{k: v for k, v in KW.items() if k != "M"})
    p0 = sdc_gym_b200.make("sdc-v0", num_envs=n2, device="cuda:0", phased=False, **kw2)
    p1 = sdc_gym_b200.make("sdc-v0", num_envs=n2, device="cuda:1", phased=True, **kw2)
    p0.reset()
    p1.reset()
    a2 = rng.uniform(0, 0.3, (n2, num_actions(5, "strictly_lower_tri")))
    q0 = p0.step_tensor(torch.as_tensor(a2, device="cuda:0"))
    q1 = p1.step_tensor(torch.as_tensor(a2, device="cuda:1"))
    assert torch.equal(q0["niter"].cpu(), q1["niter"].cpu()) and torch.equal(p0.S.cpu().view(torch.int64), p1.S.cpu().view(torch.int64))
    assert int(p1.phase_count[0]) > 0 and torch.cuda.current_device() == 0
    v1 = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=64, seed=1, device="cuda:1", **KW))
    v1.reset()
    v1.step(rng.uniform(-1, 1, (64, 5)))
    assert torch.cuda.current_device() == 0


def test_empty_batches_are_noops():
    import torch
    from sdc_gym_b200.loss import SpectralRadiusLoss

    loss = SpectralRadiusLoss(5, 1.0, "diag")
    rho = loss.spectral_radii(np.zeros(0, np.complex128), np.zeros((0, 5)))
    assert rho.shape == (0,) and rho.is_cuda
    rho, g = loss.radii_and_grads(np.zeros(0, np.complex128), np.zeros((0, 5)))
    assert rho.shape == (0,) and g.shape == (0, 5)
    env = sdc_gym_b200.make("sdc-v0", num_envs=0, seed=1, **KW)
    assert env.reset().shape == (0, 2, 5)
    obs, rew, done, infos = env.step(np.zeros((0, 5)))
    assert obs.shape == (0, 2, 5) and rew.shape == (0,) and len(infos) == 0


def test_vecnormalize_npz_round_trip_keeps_training_flag(tmp_path):
    n = 128
    rng = np.random.default_rng(0)
    e = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, **KW), gamma=0.97, clip_obs=7.0)
    e.reset()
    for _ in range(4):
        e.step(rng.uniform(-1, 1, (n, 5)))
    e.training = False
    path = str(tmp_path / "vecnormalize.pkl")  # the reference's file name (utils/utils.py:417); the content is .npz
    e.save(path)
    with np.load(path, allow_pickle=False) as z:
        assert str(z["format"]) == "sdc_gym_b200.VecNormalize" and int(z["version"]) == 1
    f = sdc_gym_b200.VecNormalize.load(path, sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, **KW))
    assert f.training is False and f.gamma == 0.97 and f.clip_obs == 7.0
    assert np.array_equal(f.obs_rms.mean.cpu().numpy(), e.obs_rms.mean.cpu().numpy())
    assert np.array_equal(f.ret_rms.var.cpu().numpy(), e.ret_rms.var.cpu().numpy()) and f.obs_rms.count == e.obs_rms.count
    assert np.array_equal(f.returns.cpu().numpy(), e.returns.cpu().numpy())
    g = sdc_gym_b200.VecNormalize.load(path, sdc_gym_b200.make("sdc-v1", num_envs=2 * n, seed=1, **KW))
    assert float(g.returns.abs().sum()) == 0.0  # per-env returns are not carried over to a different batch
    with pytest.raises(ValueError, match="M="):
        sdc_gym_b200.VecNormalize.load(path, sdc_gym_b200.make("sdc-v1", num_envs=n, seed=1, **{**KW, "M": 3}))


def test_gym_adapter_single_env_matches_oracle():
    """`gym.make('sdc-v1')` after `register_gym()` resolves to gym_adapter.SingleEnv; gym itself is optional"""
    from sdc_gym_b200.gym_adapter import make_single

    Q = collocation_matrix(5)
    env = make_single("sdc-v1", seed=3, **KW)
    u0, r0 = env.reset()
    lam = np.array([env.lam])
    u, r = exact.reset(Q, 1.0, lam)
    assert_same(u0, u[0]); assert_same(r0, r[0])
    rinit, niter = r.copy(), np.zeros(1, np.int32)
    rng = np.random.default_rng(0)
    for s in range(60):
        a = rng.uniform(-0.6, 0.0, 5)
        (ou, orr), rew, done, info = env.step(a)
        out = exact.step("sdc-v1", Q, 1.0, lam, u, r, niter, rinit, a[None])
        assert_same(ou, u[0]); assert_same(orr, r[0]); assert info["niter"] == niter[0] and info["lam"] == lam[0]
        assert done == bool(out["done"][0]) and set(info) == {"residual", "niter", "lam"}
        assert_reward_close(rew, out["reward"][0])
        if done:
            break
    assert env.M == 5 and env.restol == 1e-10 and env.prec is None
    assert isinstance(sdc_gym_b200.register_gym(), list)


def test_lazy_info_arrays_of_large_batches():
    """Batches that are pipelined in chunks leave niter / residual / lam on the device until they are read
    (lazy_info, default): same values as the eager path, fetched once, refused after the env has stepped again."""
    import sdc_gym_b200

    N, M = 40000, 5  # >= 32768: chunked pipeline
    kw = dict(num_envs=N, M=M, dt=1.0, restol=1e-10, seed=5, lambda_real_interval=[-100, 0],
              lambda_imag_interval=[-10, 0])
    lazy = sdc_gym_b200.make("sdc-v0", **kw)
    eager = sdc_gym_b200.make("sdc-v0", lazy_info=False, **kw)
    lazy.reset()
    eager.reset()
    act = np.random.default_rng(0).uniform(-1, 1, (N, M))
    o1, r1, d1, i1 = lazy.step(act)
    o2, r2, d2, i2 = eager.step(act)
    assert i1._info_fetch is not None and i2._info_fetch is None
    assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(d1, d2)
    assert np.array_equal(i1.niter, i2.niter) and i1._info_fetch is None  # fetched by the first access
    assert np.array_equal(i1.residual, i2.residual) and np.array_equal(i1.lam, i2.lam)
    d = i1[7]
    assert d["niter"] == int(i2.niter[7]) and d["lam"] == complex(i2.lam[7]) and "TimeLimit.truncated" in d
    # per-env dict access alone triggers the fetch as well
    o1, r1, d1, i1 = lazy.step(act)
    o2, r2, d2, i2 = eager.step(act)
    assert i1[3]["residual"] == i2[3]["residual"]
    # not read before the next step: gone, and said so
    _, _, _, stale = lazy.step(act)
    lazy.step(act)
    with pytest.raises(RuntimeError):
        stale.niter
    # small batches always carry their info arrays
    small = sdc_gym_b200.make("sdc-v0", **{**kw, "num_envs": 64})
    small.reset()
    _, _, _, inf = small.step(act[:64])
    assert inf._info_fetch is None and inf.niter.shape == (64,)
