"""Exact vs certified sweep mode on the benchmark workloads (device-resident steps, CUDA events):
    python tools/bench_certified.py [--envs N] [--M 5] [--steps 10]
Prints one JSON line per (workload, mode): ms per step, env-steps/s, fallback fraction, and the parity of niter / flags
between the two modes on identical inputs."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sdc_gym_b200
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--M", type=int, default=5)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    N, M = args.envs, args.M
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    KW = dict(num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0, lambda_real_interval=[-100, 0],
              lambda_imag_interval=[-10, 0], device=dev)
    x = np.diag(fixed_preconditioner("min", M, collocation_matrix(M)))
    xt = torch.as_tensor(x, device=dev)
    for workload in ("uniform", "good"):
        gen.manual_seed(1)
        if workload == "uniform":
            pool = [torch.rand((N, M), dtype=torch.float64, device=dev, generator=gen) * 2 - 1 for _ in range(4)]
        else:
            pool = [2 * (xt[None] + (torch.rand((N, M), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 0.04) - 1
                    for _ in range(4)]
        results = {}
        for mode in ("exact", "certified"):
            env = sdc_gym_b200.make("sdc-v0", sweep_mode=mode, **KW)
            env.reset()
            for k in range(3):
                env.step_tensor(pool[k % 4])
            torch.cuda.synchronize()
            fb0 = env.fallback_stats()[1]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sum_niter = 0
            for k in range(args.steps):
                env.step_tensor(pool[k % 4])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            fb = (env.fallback_stats()[1] - fb0) / (N * args.steps)
            results[mode] = (env.info_niter[:N].clone(), env.flags[:N].clone(), env.S.clone())
            flags = env.flags[:N]
            print(json.dumps({"workload": workload, "mode": mode, "M": M, "envs": N, "ms_per_step": ms,
                              "env_steps_per_s": N / ms * 1e3, "fallback_frac": fb,
                              "mean_niter": float(env.info_niter[:N].double().mean()),
                              "converged_frac": float((flags & 2).ne(0).double().mean()),
                              "err_frac": float((flags & 4).ne(0).double().mean())}), flush=True)
            del env
        a, b = results["exact"], results["certified"]
        print(json.dumps({"workload": workload, "parity_last_step": {
            "niter_equal": bool(torch.equal(a[0], b[0])), "flags_equal": bool(torch.equal(a[1], b[1])),
            "next_state_equal": bool(torch.equal(a[2], b[2]))}}), flush=True)


if __name__ == "__main__":
    main()
