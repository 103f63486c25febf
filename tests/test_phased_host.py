"""CPU pre-flight of the phased dense full solve (csrc/step_kernels.cuh, step_one PHASE): the kernel bodies, compiled
for the host by tests/host_shim and driven through the pass sequence of csrc/step_inst.cu launch_phased_dense, against
the single-pass body and the CPU oracle.  Every output must be bit-identical: the phases only regroup envs."""
import numpy as np
import pytest

from oracle import exact
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import num_actions
from tests import host_shim


def _bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.complex128:
        a = a.view(np.float64)
    return a.view(np.int64) if a.dtype == np.float64 else a


def _batches(M, n, stops, split=True, **kw):
    d = host_shim.make_desc("sdc-v0", M, seed=11, **kw)
    a = host_shim.ShimBatch(d, n)
    b = host_shim.ShimBatch(d, n, entry="shim_step_phased")
    b.phase_stops = stops
    b.phase_split = split  # True: inverse-only pass + first sweeps with the inverse reloaded; False: fused first pass
    return a, b


@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("prec_type", ["lower_tri", "strictly_lower_tri"])
@pytest.mark.parametrize("stops", [(6, 16), (1,), (2, 3, 5, 9, 30, 49)])
@pytest.mark.parametrize("split", [True, False])
def test_phased_passes_equal_the_single_pass(M, prec_type, stops, split):
    n = 150 if M <= 5 else 70  # ragged against the 128-thread blocks the shim emulates
    a, b = _batches(M, n, stops, split, prec_type=prec_type, do_scale=False, autoreset=True, strategy="residual_change")
    rng = np.random.default_rng(M)
    a.reset()
    b.reset()
    for step in range(2):
        act = rng.uniform(0, 0.3, (n, num_actions(M, prec_type)))
        oa, ob = a.step(act), b.step(act)
        for k in oa:
            assert np.array_equal(_bits(oa[k]), _bits(ob[k])), k
        for k in ("S", "lam", "resnorm", "niter", "episodes", "rng_ctr"):
            assert np.array_equal(_bits(getattr(a, k)), _bits(getattr(b, k))), k
        nit = oa["niter"]
        assert [int(c) for c in b.phase_count[: len(stops)]] == [int((nit > s).sum()) for s in stops]
        assert b.phase_count[len(stops)] == 0  # the last pass suspends nobody


def test_phased_passes_against_the_oracle():
    M, n, pt = 5, 200, "lower_tri"
    Q = collocation_matrix(M)
    rng = np.random.default_rng(3)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    act = rng.uniform(0, 0.3, (n, num_actions(M, pt)))
    _, b = _batches(M, n, (4, 12), prec_type=pt, do_scale=False, autoreset=False)
    b.reset(lam)
    out = b.step(act)
    u, r = exact.reset(Q, 1.0, lam)
    niter = np.zeros(n, np.int32)
    ref = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), act, prec_type=pt, do_scale=False)
    assert np.array_equal(out["niter"], niter)
    assert np.array_equal(_bits(out["u"]), _bits(u)) and np.array_equal(_bits(out["r"]), _bits(r))
    assert np.array_equal(_bits(out["residual"]), _bits(ref["resnorm"]))
    assert np.array_equal(_bits(out["reward"]), _bits(ref["reward"]))
    assert np.array_equal(out["conv"], ref["done"]) and np.array_equal(out["err"], ref["err"])
    assert 0 < b.phase_count[1] < b.phase_count[0] < n


@pytest.mark.parametrize("kw", [dict(prec="LU"), dict(prec_type="lower_diag"), dict(prec_type="lower_tri", cplx=True, do_scale=False),
                                dict(prec_type="lower_tri", variant=1), dict(prec_type="lower_tri", use_doubles=False)])
@pytest.mark.parametrize("split", [True, False])
def test_phased_passes_other_configurations(kw, split):
    M, n = 4, 140
    a, b = _batches(M, n, (3, 8), split, autoreset=True, **kw)
    rng = np.random.default_rng(7)
    a.reset()
    b.reset()
    A = 0 if kw.get("prec") else num_actions(M, kw["prec_type"])
    act = None
    if A:
        act = rng.uniform(0, 0.3, (n, A)) + 1j * rng.uniform(-0.05, 0.05, (n, A)) if kw.get("cplx") else rng.uniform(-1, 1, (n, A))
        if kw.get("use_doubles") is False:
            act = act.astype(np.float32).astype(np.float64)
    oa, ob = a.step(act), b.step(act)
    for k in oa:
        assert np.array_equal(_bits(oa[k]), _bits(ob[k])), k
    assert np.array_equal(_bits(a.S), _bits(b.S))
