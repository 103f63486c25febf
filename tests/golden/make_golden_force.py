"""Generates tests/golden/sdc_force_golden.npz: episodes of the reference's ``SDC_Full_Force_Env`` (``sdc-v4``,
/root/reference/sdc_gym/envs/sdc_force_env.py) on seeded lambdas and action sequences.

The class, its ``step`` and its ``reset`` are the reference's, loaded unmodified by
``oracle/ref_loader.make_reference_force_env``; the instance carries the one repair without which no non-diverging
step completes (``reward_func`` called with 4 of its 6 positional arguments, ``sdc_force_env.py:77-82`` - see the
loader).  Run in the build container only:

    OPENBLAS_NUM_THREADS=1 python tests/golden/make_golden_force.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_loader  # noqa: E402
from tests.golden.make_golden import MIN_DIAG  # noqa: E402

T = 8  # tries recorded per episode (the episode may end earlier)
CASES = [
    dict(name="m3_iter", M=3, kw=dict()),
    dict(name="m5_iter", M=5, kw=dict()),
    dict(name="m5_reschange", M=5, kw=dict(reward_iteration_only=False)),
    dict(name="m5_min", M=5, kw=dict(prec="min")),
    dict(name="m7_fast", M=7, kw=dict(reward_strategy="fast_convergence")),
]


def main():
    out = {}
    for ci, case in enumerate(CASES):
        M, n = case["M"], 24
        rng = np.random.default_rng(100 + ci)
        lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
        x = np.asarray(MIN_DIAG[M])
        # the diagonal accumulates over the tries: weights that approach the MIN diagonal after four tries, with
        # noise; every fourth env gets uniform random actions (large diagonals: slow or diverging tries)
        w = np.array([0.4, 0.3, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0])
        scaled = x[None, None, :] * w[None, :, None] + rng.uniform(0, 0.004, (n, T, M))
        actions = 2 * scaled - 1
        actions[::4] = rng.uniform(-1, 1, (n // 4 + (n % 4 > 0), T, M))
        res = np.zeros((n, T, M), np.complex128)
        diag = np.zeros((n, T, M), np.complex128)
        reward, resnorm = np.zeros((n, T)), np.zeros((n, T))
        done, valid = np.zeros((n, T), bool), np.zeros((n, T), bool)
        niter, ntries = np.zeros((n, T), np.int32), np.zeros((n, T), np.int32)
        r0 = np.zeros((n, M), np.complex128)
        for e in range(n):
            env = ref_loader.make_reference_force_env(lam=lam[e], M=M, dt=1.0, restol=1e-10,
                                                      lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
                                                      **case["kw"])
            r0[e] = env.state[0]
            for t in range(T):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")  # (fill_diagonal of a complex diagonal into the float Q_delta)
                    obs, rew, dn, info = env.step(actions[e, t])
                res[e, t], diag[e, t] = obs[0], obs[1]
                reward[e, t], done[e, t], valid[e, t] = rew, dn, True
                niter[e, t], ntries[e, t], resnorm[e, t] = info["niter"], info["ntries"], info["residual"]
                if dn:
                    break
        for k, v in dict(lam=lam, actions=actions, res=res, diag=diag, reward=reward, resnorm=resnorm, done=done,
                         valid=valid, niter=niter, ntries=ntries, r0=r0).items():
            out[f"{case['name']}/{k}"] = v
        print(case["name"], "tries", int(valid.sum()), "episodes ended", int(done.any(1).sum()), "converged-like rewards",
              int((reward > 0).sum()), "diverged", int((np.isclose(reward, -5.1)).sum()))
    np.savez_compressed(os.path.join(HERE, "sdc_force_golden.npz"), **out)


if __name__ == "__main__":
    main()
