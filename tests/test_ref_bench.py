"""CPU: the reference-timing harness (oracle/ref_bench.py) steps the unmodified reference env when it is reachable
(live tree or the copy staged in oracle/_ref) and agrees with the port on the same seeds."""
import numpy as np
import pytest

from oracle import ref_bench, ref_loader


def test_staged_reference_is_byte_identical_to_the_live_tree():
    import hashlib
    import os

    if not ref_loader.reference_available() or not os.path.isfile(ref_loader.STAGED_REF):
        pytest.skip("needs the live reference tree and a staged copy (build container after build())")
    live = os.path.join(ref_loader.REFERENCE_ROOT, "sdc_gym", "envs", "sdc_env.py")
    h = [hashlib.sha256(open(p, "rb").read()).hexdigest() for p in (live, ref_loader.STAGED_REF)]
    assert h[0] == h[1]
    assert open(ref_loader.STAGED_REF + ".sha256").read().strip() == h[0]


@pytest.mark.parametrize("kind", ["sdc-v0", "sdc-v1"])
def test_reference_loop_and_port_loop_agree(kind):
    if ref_bench.available_impl() != "reference":
        pytest.skip("no reference file reachable")
    kw = dict(M=5, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    a = ref_bench.DummyVecEnvLoop(kind, 4, "reference", seed=3, **kw)
    b = ref_bench.DummyVecEnvLoop(kind, 4, "port", seed=3, **kw)
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa, ob)
    rng = np.random.RandomState(0)
    for _ in range(60):
        act = list(rng.uniform(-1, 1, (4, 5)))
        ra, rb = a.step(act), b.step(act)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])
        for ia, ib in zip(ra[3], rb[3]):
            assert ia["niter"] == ib["niter"] and ia["lam"] == ib["lam"] and ia["residual"] == ib["residual"]
            assert ("terminal_observation" in ia) == ("terminal_observation" in ib)


def test_throughput_helpers_return_sane_numbers():
    steps, el, sum_niter, impl = ref_bench.rollout_throughput("sdc-v0", 0.2)
    assert steps >= 8 and el >= 0.2 and sum_niter >= steps and impl in ("reference", "port")
    n, el, mean_rho, impl = ref_bench.spectral_radius_throughput(0.2)
    assert n >= 256 and 0.0 < mean_rho < 1.0
    n2, _, mean_rho2, _ = ref_bench.spectral_radius_throughput(0.2, impl="port")
    assert abs(mean_rho2 - mean_rho) < 0.05  # same distribution of lambdas (same seed): close means
