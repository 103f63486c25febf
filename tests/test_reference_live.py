"""CPU, build container only: the C oracle against the unmodified reference env executed live
(skipped where /root/reference does not exist, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import exact, ref_loader
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner
from tests.helpers import assert_reward_close, assert_same

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")


@pytest.mark.parametrize("kind", ["sdc-v0", "sdc-v1"])
@pytest.mark.parametrize("M", [3, 5, 7, 9])
@pytest.mark.parametrize("prec", [None, "LU", "min", "EE", "zeros"])
def test_oracle_equals_reference_live(kind, M, prec):
    rng = np.random.default_rng(hash((kind, M, prec)) % 2**32)
    Q = collocation_matrix(M)
    Qd_fixed = fixed_preconditioner(prec, M, Q) if prec else None
    for e in range(6):
        lam = complex(rng.uniform(-100, 0), rng.uniform(-10, 0))
        env = ref_loader.make_reference_env(kind, lam=lam, M=M, dt=1.0, restol=1e-10, prec=prec,
                                            reward_iteration_only=False)
        u, r = exact.reset(Q, 1.0, [lam])
        rinit, niter = r.copy(), np.zeros(1, np.int32)
        for s in range(50 if kind == "sdc-v1" else 1):
            a = rng.uniform(-1, 1, M)
            _, rew, done, info = env.step(a.copy())
            out = exact.step(kind, Q, 1.0, [lam], u, r, niter, rinit, None if prec else a[None],
                             prec_type="fixed" if prec else "diag", Qd_fixed=Qd_fixed,
                             reward_strategy="residual_change")
            assert_same(u[0], env.state[0]); assert_same(r[0], env.state[1])
            assert info["niter"] == niter[0]
            assert_same(out["resnorm"][0], info["residual"])
            assert_reward_close(out["reward"][0], rew)
            if kind == "sdc-v1":
                assert bool(done) == bool(out["done"][0])
            if done:
                break


def test_fixed_preconditioners_equal_reference():
    mod = ref_loader.load_reference_envs()
    for M in (2, 3, 4, 5, 6, 7, 9):
        for prec in ("LU", "min", "EE", "zeros"):
            env = mod.SDC_Full_Env(M=M, dt=1.0, restol=1e-10, prec=prec)
            assert_same(np.asarray(env._get_prec(None), dtype=np.float64), fixed_preconditioner(prec, M), f"{prec} M={M}")
