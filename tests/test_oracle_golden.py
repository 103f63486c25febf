"""CPU: the rounding-exact C oracle (oracle/sdc_exact.c) replays every golden case bit-for-bit.

The golden vectors were produced by the unmodified reference env (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from oracle import exact
from tests.helpers import assert_reward_close, assert_same, case_arrays, case_ids, case_meta, golden


def replay_with_oracle(name):
    meta, g = case_meta(name), case_arrays(name)
    M, kind, n = meta["M"], meta["kind"], meta["n"]
    Q = g["Q"]
    u, r = exact.reset(Q, meta["dt"], g["lam"], variant=0)
    assert_same(u, g["u0"], f"{name} reset u")
    assert_same(r, g["r0"], f"{name} reset r")
    rinit = r.copy()
    niter = np.zeros(n, np.int32)
    Qd_fixed = None
    if meta["prec"] is not None:
        from sdc_gym_b200.precond import fixed_preconditioner
        Qd_fixed = fixed_preconditioner(meta["prec"], M, Q)
    A = exact.num_actions(M, meta["prec_type"])
    steps_max = g["u"].shape[1]
    old_states = None
    if meta["collect"] and kind == "sdc-v0":
        old_states = np.zeros((n, 2 * M, 50), np.complex128)
        old_states[:, :, 0] = np.concatenate((u, r), axis=1)
    alive = np.ones(n, bool)
    for s in range(steps_max):
        act = g["actions"][:, s, :A] if A else None
        out = exact.step(kind, Q, meta["dt"], g["lam"], u, r, niter, rinit, act, prec_type=meta["prec_type"],
                         Qd_fixed=Qd_fixed, do_scale=meta["do_scale"], use_doubles=meta.get("use_doubles", True),
                         reward_strategy=meta["strategy"],
                         step_penalty=meta["step_penalty"], residual_weight=meta["residual_weight"],
                         norm_factor=meta["norm_factor"], restol=meta["restol"], variant=0, collect_states=old_states)
        live = alive & (s < g["nsteps"])
        assert_same(u[live], g["u"][live, s], f"{name} step {s} u")
        assert_same(r[live], g["r"][live, s], f"{name} step {s} r")
        assert_same(out["resnorm"][live], g["residual"][live, s], f"{name} step {s} residual")
        assert np.array_equal(niter[live], g["niter"][live, s]), f"{name} step {s} niter"
        assert_reward_close(out["reward"][live], g["reward"][live, s], f"{name} step {s}")
        if kind == "sdc-v1":
            assert np.array_equal(out["done"][live], g["done"][live, s]), f"{name} step {s} done"
    if old_states is not None:
        assert_same(old_states, g["old_states"], f"{name} old_states")


@pytest.mark.parametrize("name", case_ids())
def test_oracle_replays_golden(name):
    replay_with_oracle(name)


def test_manifest_records_blas_build():
    manifest, _ = golden()
    assert manifest["blas_variant"] == 0 and manifest["openblas_core"] == "SkylakeX"
    assert len(manifest["cases"]) >= 100
