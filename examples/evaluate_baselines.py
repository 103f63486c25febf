#!/usr/bin/env python
"""The evaluation loop of the reference (`rl_playground.test_model` / `run_tests`, rl_playground.py:89-249) on the
batched env: fixed preconditioners LU and MIN (and optionally a constant diagonal "policy") on freshly drawn lambdas.

    python examples/evaluate_baselines.py --envname sdc-v0 --M 5 --num_envs 100000 --tests 5

Prints the reference's summary line per preconditioner:
    LU  -- Mean number of iterations and success rate: 17.13, 100.0 %
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdc_gym_b200  # noqa: E402


def test_model(env, ntests, name, predict=None):
    """rl_playground.test_model, statement for statement, on the vectorised env."""
    mean_niter, nsucc = 0, 0
    num_envs = env.num_envs
    for _ in range(ntests):
        obs = env.reset()
        done = np.zeros(num_envs, bool)
        if env.envs[0].prec is not None:
            action = [np.empty(env.action_space.shape, dtype=env.action_space.dtype)] * 0 or None
        while not done.all():
            if env.envs[0].prec is None:
                action = predict(obs)
            obs, rewards, done, info = env.step(action)
        # `info` keeps the finished episode (the env itself has already been reset): arrays instead of a dict loop
        ok = (info.niter < 50) & (info.residual < env.restol)
        nsucc += int(ok.sum())
        mean_niter += int(info.niter[ok].sum())
    mean_niter = mean_niter / nsucc if nsucc else 666
    print(f"{name:<3} -- Mean number of iterations and success rate: {mean_niter:4.2f}, "
          f"{nsucc / (ntests * num_envs) * 100} %")
    return mean_niter, nsucc / (ntests * num_envs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envname", default="sdc-v0")
    ap.add_argument("--M", type=int, default=5)
    ap.add_argument("--num_envs", type=int, default=100000)
    ap.add_argument("--tests", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    kw = dict(M=args.M, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[0, 0],
              seed=args.seed)
    out = {}
    for prec in ("LU", "min"):
        env = sdc_gym_b200.make(args.envname, num_envs=args.num_envs, prec=prec, **kw)
        out[prec] = test_model(env, args.tests, prec.upper())
    # a constant diagonal policy (the MIN diagonal expressed as actions in [-1, 1])
    x = np.diag(sdc_gym_b200.fixed_preconditioner("min", args.M))
    env = sdc_gym_b200.make(args.envname, num_envs=args.num_envs, **kw)
    out["policy"] = test_model(env, args.tests, "RL", predict=lambda obs: np.tile(2 * x - 1, (args.num_envs, 1)))
    return out


if __name__ == "__main__":
    main()
