"""CPU: the numpy port (oracle/sdc_port.py, the thing bench.py times as the CPU baseline) replays the golden
vectors of the unmodified reference env.  Bit-equality holds when this host's numpy/OpenBLAS dispatch is the one
the vectors were made with (SkylakeX core); otherwise the comparison drops to 1e-9 relative on states."""
import numpy as np
import pytest

from oracle import sdc_port
from sdc_gym_b200.vec_env import detect_blas_variant
from tests.helpers import assert_reward_close, assert_same, case_arrays, case_ids, case_meta

EXACT = detect_blas_variant() == 0


def _close(a, b, what):
    if EXACT:
        assert_same(a, b, what)
    else:
        assert np.allclose(a, b, rtol=1e-9, atol=1e-13), what


@pytest.mark.parametrize("name", [n for n in case_ids() if ("M5" in n or "M3" in n)])
def test_port_replays_golden(name):
    meta, g = case_meta(name), case_arrays(name)
    kind, M = meta["kind"], meta["M"]
    A = g["actions"].shape[2]
    for e in range(meta["n"]):
        env = sdc_port.ENV_CLASSES[kind](
            M=M, dt=meta["dt"], restol=meta["restol"], prec=meta["prec"], reward_iteration_only=None,
            reward_strategy=meta["strategy"], norm_factor=meta["norm_factor"], do_scale=meta["do_scale"], use_doubles=meta.get("use_doubles", True),
            free_action_space=meta["cplx"], collect_states=meta["collect"], step_penalty=meta["step_penalty"],
            residual_weight=meta["residual_weight"], prec_type=meta["prec_type"] if meta["prec"] is None else "diag")
        env.niter = 0
        env.set_lambda(g["lam"][e])
        _close(env.state[1], g["r0"][e], f"{name} r0")
        for s in range(int(g["nsteps"][e])):
            a = None if meta["prec"] is not None else g["actions"][e, s, :A].copy()
            _, rew, done, info = env.step(a)
            _close(env.state[0], g["u"][e, s], f"{name} env {e} step {s} u")
            _close(env.state[1], g["r"][e, s], f"{name} env {e} step {s} r")
            if EXACT:
                assert info["niter"] == g["niter"][e, s]
                assert_same(info["residual"], g["residual"][e, s])
                assert_reward_close(rew, g["reward"][e, s])
                assert bool(done) == bool(g["done"][e, s])
        if meta["collect"] and EXACT:
            assert_same(env.old_states, g["old_states"][e])


def test_port_vec_env_autoreset_protocol():
    vec = sdc_port.PortDummyVecEnv("sdc-v1", 3, seed=1, M=3, dt=1.0, restol=1e-10,
                                   lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    obs = vec.reset()
    assert obs.shape == (3, 2, 3) and obs.dtype == np.complex128
    rng = np.random.RandomState(0)
    seen_done = False
    for _ in range(60):
        lam_before = [e.lam for e in vec.envs]
        obs, rew, done, infos = vec.step(list(rng.uniform(-1, 1, (3, 3))))
        for i in range(3):
            if done[i]:
                seen_done = True
                assert "terminal_observation" in infos[i] and infos[i]["lam"] == lam_before[i]
                assert vec.envs[i].niter == 0 and np.all(obs[i, 0] == 1)
    assert seen_done
    steps, el, _ = sdc_port.rollout_throughput("sdc-v0", 0.2, num_envs=2, M=3)
    assert steps > 0 and el > 0
