"""GPU: the phased full solve of dense Q_delta (csrc/step_kernels.cuh, step_one PHASE; include/sdcgym.h
sdcgym_state.phase_*) against the single-launch kernel of the same library and against the CPU oracle.  The phases only
regroup envs into warps: every env runs the same sweep sequence, so EVERY output plane is bit-identical."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import sdc_gym_b200
from oracle import exact
from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import num_actions

pytestmark = pytest.mark.gpu
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          blas_variant=_lib.BLAS_SKYLAKEX)
STATE = ("S", "lam", "resnorm", "niter", "episodes", "rng_ctr")


def _pair(M, n, **kw):
    a = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=5, phased=False, **{**KW, **kw})
    b = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=5, phased=True, **{**KW, **kw})
    assert not a.phased and b.phased
    a.reset()
    b.reset()
    return a, b


def _same_step(a, b, act):
    t = torch.as_tensor(act, device=a.device)
    oa = {k: v.clone() for k, v in a.step_tensor(t).items()}
    ob = {k: v.clone() for k, v in b.step_tensor(t).items()}
    assert oa.keys() == ob.keys()
    for k in oa:
        va, vb = oa[k], ob[k]
        if va.is_floating_point():  # bit patterns (NaN rewards / residuals of diverged envs included)
            va, vb = va.view(torch.int64), vb.view(torch.int64)
        assert torch.equal(va, vb), k
    for k in STATE:
        va, vb = getattr(a, k), getattr(b, k)
        if va.is_floating_point():
            va, vb = va.view(torch.int64), vb.view(torch.int64)
        assert torch.equal(va, vb), k
    return oa


@pytest.mark.parametrize("M", [2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("prec_type", ["lower_tri", "strictly_lower_tri"])
def test_phased_solve_is_bit_identical_to_the_single_launch(M, prec_type):
    n = 20000 + 37  # ragged: the last block and the last warp of every list are partial
    a, b = _pair(M, n, prec_type=prec_type, do_scale=False)
    rng = np.random.default_rng(10 * M)
    for s in range(3):
        act = rng.uniform(0, 0.3, (n, num_actions(M, prec_type)))
        out = _same_step(a, b, act)
        c = b.phase_count.cpu().numpy()
        # default plan: one hand-over (warps that have thinned out after 5 sweeps), the second pass runs every env it
        # gets to its end
        assert 0 <= c[0] < n and c[1] == 0
        if prec_type == "strictly_lower_tri":
            assert c[0] > 0


@pytest.mark.parametrize("kw", [
    dict(prec_type="lower_diag"), dict(prec="LU"), dict(prec_type="lower_tri", reward_iteration_only=False),
    dict(prec_type="lower_tri", reward_strategy="smooth_fast_convergence"), dict(prec_type="lower_tri", autoreset=False),
    dict(prec_type="lower_tri", free_action_space=True, do_scale=False),
    dict(prec_type="strictly_lower_tri", blas_variant=_lib.BLAS_HASWELL), dict(prec_type="lower_tri", use_doubles=False),
])
def test_phased_solve_configurations(kw):
    M, n = 5, 16384
    a, b = _pair(M, n, **kw)
    rng = np.random.default_rng(2)
    A = a.n_act if kw.get("prec") is None else 0
    for s in range(2):
        if kw.get("free_action_space"):
            act = rng.uniform(0, 0.3, (n, A)) + 1j * rng.uniform(-0.05, 0.05, (n, A))
        else:
            act = rng.uniform(-1, 1, (n, max(A, 1)))
        if kw.get("use_doubles") is False:
            act = act.astype(np.float32).astype(np.float64)
        _same_step(a, b, act)
        if kw.get("autoreset") is False:
            a.reset()
            b.reset()


def test_phased_solve_against_cpu_oracle():
    M, n, pt = 5, 16384, "lower_tri"
    Q = collocation_matrix(M)
    rng = np.random.default_rng(0)
    lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
    act = rng.uniform(0, 0.3, (n, num_actions(M, pt)))
    env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, prec_type=pt, do_scale=False, autoreset=False, phased=True, **KW)
    assert env.phased
    env.reset(lam=lam)
    _, rew, done, infos = env.step(act)
    u, r = exact.reset(Q, 1.0, lam)
    niter = np.zeros(n, np.int32)
    out = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), act, prec_type=pt, do_scale=False)
    assert np.array_equal(infos.niter, niter)
    assert np.array_equal((infos.flags & 2) != 0, out["done"]) and np.array_equal((infos.flags & 4) != 0, out["err"])
    snap = env._snapshot()
    assert np.array_equal(snap["obs"][:, 0].view(np.int64), u.view(np.int64))
    assert np.array_equal(snap["obs"][:, 1].view(np.int64), r.view(np.int64))
    assert np.array_equal(np.asarray(infos.residual).view(np.int64), out["resnorm"].view(np.int64))
    assert np.array_equal(np.asarray(rew).view(np.int64), out["reward"].view(np.int64))


def test_phased_solve_through_the_chunked_host_step():
    # numpy in / numpy out: the batch is pipelined in chunks, every chunk with its own slice of the work buffers
    M, n, pt = 5, 150000, "lower_tri"
    a, b = _pair(M, n, prec_type=pt, do_scale=False)
    rng = np.random.default_rng(1)
    for s in range(2):
        act = rng.uniform(0, 0.3, (n, num_actions(M, pt)))
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(oa.view(np.int64), ob.view(np.int64))
        assert np.array_equal(np.asarray(ra).view(np.int64), np.asarray(rb).view(np.int64)) and np.array_equal(da, db)
        assert np.array_equal(ia.niter, ib.niter) and np.array_equal(ia.flags, ib.flags)
        assert np.array_equal(np.asarray(ia.residual).view(np.int64), np.asarray(ib.residual).view(np.int64))


def test_default_times_both_launch_sequences_and_keeps_one():
    M, n, pt = 4, 20000, "strictly_lower_tri"
    a = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=5, phased=False, prec_type=pt, do_scale=False, **KW)
    b = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, seed=5, prec_type=pt, do_scale=False, **KW)  # phased=None
    assert b.phased and b._phase_auto and b.phase_timings is None
    a.reset()
    b.reset()
    rng = np.random.default_rng(3)
    for s in range(12):
        _same_step(a, b, rng.uniform(0, 0.3, (n, num_actions(M, pt))))
        torch.cuda.synchronize()
    # 2 x 3 launches measured, read back at the next step
    assert b.phase_timings is not None and all(t > 0 for t in b.phase_timings)
    assert b._phase_use == (b.phase_timings[0] < b.phase_timings[1]) and b._phase_trial is None
    # the host step follows the choice
    act = rng.uniform(0, 0.3, (n, num_actions(M, pt)))
    oa, ra, da, ia = a.step(act)
    ob, rb, db, ib = b.step(act)
    assert np.array_equal(oa.view(np.int64), ob.view(np.int64)) and np.array_equal(ia.niter, ib.niter)


def test_small_batches_and_team_sizes_keep_the_single_launch():
    e = sdc_gym_b200.make("sdc-v0", num_envs=64, M=5, prec_type="lower_tri", **KW)
    assert not e.phased
    e = sdc_gym_b200.make("sdc-v0", num_envs=64, M=5, prec_type="lower_tri", phased=True, **KW)
    assert e.phased  # (explicitly asked for: any batch size)
    e = sdc_gym_b200.make("sdc-v0", num_envs=20000, M=9, prec_type="lower_tri", **KW)
    assert not e.phased
    e = sdc_gym_b200.make("sdc-v0", num_envs=20000, M=5, prec_type="diag", **KW)
    assert not e.phased
    e = sdc_gym_b200.make("sdc-v1", num_envs=20000, M=5, prec_type="lower_tri", **KW)
    assert not e.phased
    e = sdc_gym_b200.make("sdc-v0", num_envs=20000, M=5, prec_type="lower_tri", **KW)
    assert e.phased


_OTHER_STOPS = r"""
import numpy as np, torch, sdc_gym_b200
from sdc_gym_b200.precond import num_actions
kw = dict(M=4, prec_type="lower_tri", do_scale=False, dt=1.0, restol=1e-10, seed=5,
          lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
n = 17000
a = sdc_gym_b200.make("sdc-v0", num_envs=n, phased=False, **kw)
b = sdc_gym_b200.make("sdc-v0", num_envs=n, phased=True, **kw)
a.reset(); b.reset()
act = torch.as_tensor(np.random.default_rng(0).uniform(0, 0.3, (n, num_actions(4, "lower_tri"))), device=a.device)
oa = {k: v.clone() for k, v in a.step_tensor(act).items()}
ob = {k: v.clone() for k, v in b.step_tensor(act).items()}
for k in oa:
    va, vb = (oa[k].view(torch.int64), ob[k].view(torch.int64)) if oa[k].is_floating_point() else (oa[k], ob[k])
    assert torch.equal(va, vb), k
assert torch.equal(a.S.view(torch.int64), b.S.view(torch.int64))
c = b.phase_count.cpu().numpy(); nit = oa["niter"].cpu().numpy()
import os
if "SDCGYM_PHASE_LANES" not in os.environ and "SDCGYM_PHASE_STOPS" in os.environ:  # fixed sweep counts only: the list lengths are known
    stops = [int(x) for x in os.environ["SDCGYM_PHASE_STOPS"].split(",")]
    assert [int(x) for x in c[:len(stops)]] == [int((nit > s).sum()) for s in stops], (c, stops)
else:
    assert c[0] > 0
print("ok")
"""


@pytest.mark.parametrize("plan", [dict(SDCGYM_PHASE_STOPS="1"), dict(SDCGYM_PHASE_STOPS="2,3,4,5,49"),
                                  dict(SDCGYM_PHASE_STOPS="10,20,30,40,45,49"), dict(SDCGYM_PHASE_STOPS="50"),
                                  dict(SDCGYM_PHASE_LANES="32"), dict(SDCGYM_PHASE_LANES="8,8,8,8,8,8"),
                                  dict(SDCGYM_PHASE_LANES="28,20,12"), dict(SDCGYM_PHASE_STOPS="4,12", SDCGYM_PHASE_LANES="16,16"),
                                  dict(SDCGYM_PHASE_SPLIT="0"), dict(SDCGYM_PHASE_SPLIT="0", SDCGYM_PHASE_STOPS="3,9")])
def test_other_hand_over_rules(plan):
    env = {k: v for k, v in os.environ.items() if k not in ("SDCGYM_PHASE_STOPS", "SDCGYM_PHASE_LANES", "SDCGYM_PHASE_SPLIT")}
    env.update(plan)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _OTHER_STOPS], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.parametrize("kind,kw", [("sdc-v0", dict(prec_type="strictly_lower_tri", do_scale=False, phased=True)),
                                     ("sdc-v0", dict(prec_type="lower_tri", do_scale=False)),  # phased=None: no timing inside a capture
                                     ("sdc-v0", dict()), ("sdc-v1", dict())])
def test_device_steps_can_be_captured_in_a_cuda_graph(kind, kw):
    """the device-resident step is stream-ordered launches (and one memset for the phased solve): capturable; a replayed
    graph of three steps equals three eager steps bit for bit"""
    M, n = 5, 1 << 16
    mk = lambda: sdc_gym_b200.make(kind, num_envs=n, M=M, seed=8, **{**KW, **kw})
    a, b = mk(), mk()
    a.reset()
    b.reset()
    A = a.n_act
    hi = 0.3 if kw.get("do_scale") is False else 1.0
    gen = torch.Generator(device=a.device)
    gen.manual_seed(4)
    acts = [torch.rand((n, A), dtype=torch.float64, device=a.device, generator=gen) * hi for _ in range(4)]
    for e in (a, b):  # first launches configure the kernels (shared-memory opt-in): outside the capture
        e.step_tensor(acts[3])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for k in range(3):
                b.step_tensor(acts[k])
    torch.cuda.current_stream().wait_stream(side)
    for k in range(3):
        out = a.step_tensor(acts[k])
    g.replay()
    torch.cuda.synchronize()
    for name in STATE:
        va, vb = getattr(a, name), getattr(b, name)
        if va.is_floating_point():
            va, vb = va.view(torch.int64), vb.view(torch.int64)
        assert torch.equal(va, vb), name
    assert torch.equal(a.info_niter, b.info_niter) and torch.equal(a.reward.view(torch.int64), b.reward.view(torch.int64))


@pytest.mark.parametrize("n", [1, 33, 129, 1000])
def test_tiny_batches_when_asked_for(n):
    a, b = _pair(5, n, prec_type="strictly_lower_tri", do_scale=False)
    rng = np.random.default_rng(n)
    for s in range(3):
        _same_step(a, b, rng.uniform(0, 0.3, (n, num_actions(5, "strictly_lower_tri"))))


_TMEM_CHECK = r"""
import numpy as np, torch, sys
import sdc_gym_b200
from oracle import exact
from sdc_gym_b200.collocation import collocation_matrix
M, n = int(sys.argv[1]), 20000 + 11
Q = collocation_matrix(M)
rng = np.random.default_rng(M)
lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
act = rng.uniform(-1, 1, (n, M))
env = sdc_gym_b200.make("sdc-v0", num_envs=n, M=M, dt=1.0, restol=1e-10, autoreset=False,
                        lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
env.reset(lam=lam)
_, rew, done, infos = env.step(act)
u, r = exact.reset(Q, 1.0, lam)
niter = np.zeros(n, np.int32)
out = exact.step("sdc-v0", Q, 1.0, lam, u, r, niter, r.copy(), act)
snap = env._snapshot()
assert np.array_equal(snap["obs"][:, 0].view(np.int64), u.view(np.int64))
assert np.array_equal(snap["obs"][:, 1].view(np.int64), r.view(np.int64))
assert np.array_equal(infos.niter, niter)
assert np.array_equal(np.asarray(infos.residual).view(np.int64), out["resnorm"].view(np.int64))
print("ok")
"""


@pytest.mark.parametrize("M", [2, 3, 4, 5])
def test_system_matrix_in_tensor_memory_variant_equals_oracle(M):
    """SDCGYM_TMEM=1 (experiment switch, csrc/step_kernels.cuh step_tmem_kernel): the diagonal full solve with every
    env's C in tensor memory (tcgen05.alloc / st / ld) instead of shared memory - a different store for the same
    numbers, so the oracle comparison is bit for bit.  Own process (the switch is read once) with a time limit."""
    env = dict(os.environ, SDCGYM_TMEM="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _TMEM_CHECK, str(M)], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
