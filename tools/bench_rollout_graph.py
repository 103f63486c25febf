"""Experiment driver (GPU): rollout collection (normalised sdc-v1, device policy, RolloutBuffer, GAE) eagerly and as one
CUDA graph replay (sdc_gym_b200.rollout.GraphedRollout), over the batch size.  One JSON line per size."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
from sdc_gym_b200.rollout import GraphedRollout, collect_rollouts

T, M = 8, 5
KW = dict(M=M, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0,
          output="torch", reward_iteration_only=False)
w = torch.linspace(-1.0, 1.0, 4 * M, dtype=torch.float64, device="cuda")


def policy(obs_planes):
    s = torch.tanh((obs_planes * w[:, None]).sum(0))
    a = torch.stack([torch.tanh(s * (k + 1) * 0.7) for k in range(M)], dim=1) * 0.9
    return a, s * 0.1, None


for n in (8, 64, 1024, 16384, 131072, 1 << 20):
    rec = {"envs": n, "n_steps": T}
    for mode in ("eager", "graph"):
        env = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=n, **KW))
        env.reset()
        gr = GraphedRollout(env, policy, T, warmup=2) if mode == "graph" else None
        buf = None

        def once():
            global buf
            if gr is not None:
                return gr.collect()
            buf = collect_rollouts(env, policy, T, buffer=buf)

        for _ in range(4):
            once()
        torch.cuda.synchronize()
        reps = 20 if n <= 131072 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        rec[mode + "_ms_per_rollout"] = round(ms, 4)
        rec[mode + "_env_steps_per_s"] = round(n * T / ms * 1e3, 1)
        del env
    rec["speedup"] = round(rec["eager_ms_per_rollout"] / rec["graph_ms_per_rollout"], 2)
    print(json.dumps(rec), flush=True)
