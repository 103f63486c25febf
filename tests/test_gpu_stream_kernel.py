"""GPU: the persistent, bulk-async-pipelined sdc-v1 kernel (csrc/stream_kernels.cuh) - taken for batches of at least
296 full tiles of 256 envs - against the CPU oracle, bit for bit, including the tail that does not fill a tile, complex
actions, fixed preconditioners and the fused auto-reset."""
import numpy as np
import pytest
import torch

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner
from tests.helpers import assert_reward_close, assert_same

pytestmark = pytest.mark.gpu
N = 148 * 2 * 256 + 77  # 296 full tiles of 256 envs (the streaming kernel's threshold) + a tail for the plain kernel
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)


@pytest.mark.parametrize("M", [3, 5, 7])
@pytest.mark.parametrize("variant", ["real", "complex", "min", "residual_change"])  # (the last one: with the staged norm_init plane)
def test_streaming_step_equals_oracle(M, variant):
    Q = collocation_matrix(M)
    rng = np.random.default_rng(M)
    lam = rng.uniform(-100, 0, N) + 1j * rng.uniform(-10, 0, N)
    kw, okw = {}, {}
    x = np.diag(fixed_preconditioner("min", M, Q))
    if not x.any():
        x = np.full(M, 0.2)
    if variant == "complex":
        kw = dict(free_action_space=True, do_scale=False)
        okw = dict(do_scale=False)
    elif variant == "min":
        kw = dict(prec="min")
        okw = dict(prec_type="fixed", Qd_fixed=fixed_preconditioner("min", M, Q))
    elif variant == "residual_change":
        kw = dict(reward_iteration_only=False)
        okw = dict(reward_strategy="residual_change")
    env = sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, autoreset=False, **KW, **kw)
    env.reset(lam=lam)
    u, r = exact.reset(Q, 1.0, lam)
    rinit, niter = r.copy(), np.zeros(N, np.int32)
    for step in range(4):
        act = x[None] + rng.uniform(-0.05, 0.05, (N, M))
        if variant == "complex":
            act = act + 1j * rng.uniform(-0.02, 0.02, (N, M))
        else:
            act = 2 * act - 1
        t = None if variant == "min" else torch.as_tensor(act, device=env.device)
        out = env.step_tensor(t)
        o = exact.step("sdc-v1", Q, 1.0, lam, u, r, niter, rinit, None if variant == "min" else act, **okw)
        snap = env._snapshot()
        assert_same(snap["obs"][:, 0], u, f"u, step {step}")
        assert_same(snap["obs"][:, 1], r, f"r, step {step}")
        assert np.array_equal(out["niter"].cpu().numpy(), niter)
        assert_same(out["residual"].cpu().numpy(), o["resnorm"])
        f = out["flags"].cpu().numpy()
        assert np.array_equal((f & 1) != 0, o["done"]) and np.array_equal((f & 4) != 0, o["err"])
        assert_reward_close(out["reward"].cpu().numpy(), o["reward"])


def test_streaming_step_with_autoreset_matches_small_batches():
    """The fused auto-reset (episode counters, Philox counters, new lambda, initial state) through the streaming kernel:
    the same envs stepped as one large batch and as slices small enough for the plain kernel give the same bits."""
    M, steps = 5, 8
    big = sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, seed=4, **KW)
    parts = [sdc_gym_b200.make("sdc-v1", num_envs=cnt, M=M, seed=4, env_offset=off, **KW)
             for off, cnt in ((0, 20000), (20000, N - 20000 - 3000), (N - 3000, 3000))]
    big.reset()
    for p in parts:
        p.reset()
    rng = np.random.default_rng(9)
    x = np.diag(fixed_preconditioner("min", M))
    for s in range(steps):
        a = 2 * (x[None] + rng.uniform(-0.03, 0.03, (N, M))) - 1
        a[::2] = rng.uniform(-1, 1, (N - N // 2, M))  # half of the envs with random actions: some diverge and restart
        act = torch.as_tensor(a, device=big.device)
        big.step_tensor(act)
        off = 0
        for p in parts:
            p.step_tensor(act[off:off + p.num_envs])
            for name in ("S", "lam"):
                assert torch.equal(getattr(big, name)[:, off:off + p.num_envs], getattr(p, name)[:, :p.num_envs]), (name, s)
            for name in ("resnorm", "niter", "episodes", "rng_ctr", "reward", "flags", "info_niter"):
                assert torch.equal(getattr(big, name)[off:off + p.num_envs], getattr(p, name)[:p.num_envs]), (name, s)
            off += p.num_envs
    assert int(big.episodes[:N].max()) > 1  # episodes ended and restarted on the way
