"""Replay of a golden case through a backend (host shim of the kernel templates, or the CUDA library)."""
from __future__ import annotations

import numpy as np

from tests.helpers import assert_reward_close, assert_same, case_arrays, case_meta


def replay_case(name, make_backend):
    """make_backend(meta, g) -> object with reset(lam) -> (u, r) and step(actions) -> dict (see ShimBatch.step)."""
    meta, g = case_meta(name), case_arrays(name)
    kind, n = meta["kind"], meta["n"]
    be = make_backend(meta, g)
    u, r = be.reset(g["lam"])
    assert_same(u, g["u0"], f"{name} reset u")
    assert_same(r, g["r0"], f"{name} reset r")
    A = be.n_act
    steps_max = g["u"].shape[1]
    for s in range(steps_max):
        act = g["actions"][:, s, :A] if A else None
        out = be.step(act)
        live = s < g["nsteps"]
        assert_same(out["u"][live], g["u"][live, s], f"{name} step {s} u")
        assert_same(out["r"][live], g["r"][live, s], f"{name} step {s} r")
        assert_same(out["residual"][live], g["residual"][live, s], f"{name} step {s} residual")
        assert np.array_equal(out["niter"][live], g["niter"][live, s]), f"{name} step {s} niter"
        assert_reward_close(out["reward"][live], g["reward"][live, s], f"{name} step {s}")
        assert np.array_equal(out["done"][live], g["done"][live, s]), f"{name} step {s} done"
        restol = meta["restol"]
        assert np.array_equal(out["conv"][live], g["residual"][live, s] < restol) or kind == "sdc-v0"
    if meta["collect"]:
        got = be.old_states_host()
        for e in range(n):
            # the replay keeps stepping envs that already finished: only the columns the reference wrote count
            ncol = 50 if kind == "sdc-v0" else min(int(g["nsteps"][e]) + 1, 50)
            assert_same(got[e][:, :ncol], g["old_states"][e][:, :ncol], f"{name} old_states env {e}")
