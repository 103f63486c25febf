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


@pytest.mark.skipif(not ref_loader.force_reference_available(), reason="reference sdc_force_env.py not present")
def test_force_env_fixture_is_what_the_reference_produces():
    """tests/golden/sdc_force_golden.npz (the parity anchor of ``sdc-v4``) against ``SDC_Full_Force_Env`` run live -
    and the unrepaired class does fail the way DESIGN.md says (``sdc_force_env.py:77-82``)."""
    import os
    import warnings
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "sdc_force_golden.npz"))
    for name, M, kw in (("m3_iter", 3, {}), ("m5_reschange", 5, dict(reward_iteration_only=False))):
        g = {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(name + "/")}
        for e in range(0, 24, 5):
            env = ref_loader.make_reference_force_env(lam=g["lam"][e], M=M, dt=1.0, restol=1e-10,
                                                      lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], **kw)
            assert_same(env.state[0], g["r0"][e])
            for t in range(g["actions"].shape[1]):
                if not g["valid"][e, t]:
                    break
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    obs, rew, done, info = env.step(g["actions"][e, t])
                assert_same(obs[0], g["res"][e, t]); assert_same(obs[1], g["diag"][e, t])
                assert (rew, bool(done), info["niter"], info["ntries"]) == (g["reward"][e, t], bool(g["done"][e, t]),
                                                                           g["niter"][e, t], g["ntries"][e, t])
    raw = ref_loader.load_reference_force_env().SDC_Full_Force_Env(M=3, dt=1.0, restol=1e-10)
    raw.reset()
    with pytest.raises(TypeError), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # the MIN diagonal: the solve does not diverge -> reward_func(4 of its 6 arguments)
        raw.step(2 * np.array([0.3203856825077055, 0.1399680686269595, 0.3716708461097372]) - 1)
