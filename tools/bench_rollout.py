"""Rollout-side benchmark (BASELINE config 5: sdc-v1 rollout collection with device VecNormalize): raw sdc-v1 step,
normalised step, and full collect_rollouts (random policy, RolloutBuffer, GAE).  One JSON object per line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
from sdc_gym_b200.rollout import collect_rollouts

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--M", type=int, default=5)
ap.add_argument("--only-rollout", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
N, T, M = args.envs, args.steps, args.M
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0)
gen = torch.Generator(device=dev); gen.manual_seed(0)


def timed(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if args.only_rollout:
    env = None
else:
    env = sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, reward_iteration_only=False, **KW)
    env.reset()
if env is not None:
    acts = [torch.rand((N, M), dtype=torch.float64, device=dev, generator=gen) * 2 - 1 for _ in range(2)]
    k = [0]
    def f1():
        k[0] += 1
        env.step_tensor(acts[k[0] % 2])
    ms = timed(f1, 50)
    print(json.dumps({"config": "sdc-v1 step (reward residual_change)", "M": M, "envs": N, "ms_per_step": ms,
                      "env_steps_per_s": N / ms * 1e3}), flush=True)
    env0 = sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, **KW)  # default reward: iteration_only
    env0.reset()
    def f0():
        k[0] += 1
        env0.step_tensor(acts[k[0] % 2])
    ms = timed(f0, 50)
    print(json.dumps({"config": "sdc-v1 step (reward iteration_only)", "M": M, "envs": N, "ms_per_step": ms,
                      "env_steps_per_s": N / ms * 1e3, "algorithmic_GBps": N * (8 * M + 64 * M + 49) / ms / 1e6}), flush=True)
    del env0
    vn = sdc_gym_b200.VecNormalize(env, norm_obs=True, norm_reward=True)
    vn.reset()
    def f2():
        k[0] += 1
        vn.step_tensor(acts[k[0] % 2])
    ms = timed(f2, 50)
    print(json.dumps({"config": "sdc-v1 step + device VecNormalize(norm_obs, norm_reward)", "M": M, "envs": N,
                      "ms_per_step": ms, "env_steps_per_s": N / ms * 1e3}), flush=True)
    del vn, env
env_r = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, reward_iteration_only=False, **KW))
env_r.reset()
def policy(obs_planes):
    a = torch.empty((N, M), dtype=torch.float64, device=dev).uniform_(-1.0, 1.0, generator=gen)
    return a, obs_planes[0], None
buf = collect_rollouts(env_r, policy, T)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    buf = collect_rollouts(env_r, policy, T, buffer=buf)
e1.record()
torch.cuda.synchronize()
el = e0.elapsed_time(e1) * 1e-3 / 4
print(json.dumps({"config": "sdc-v1 rollout collection (VecNormalize + RolloutBuffer + GAE, random policy)", "M": M,
                  "envs": N, "n_steps": T, "s_per_rollout": el, "env_steps_per_s": N * T / el,
                  "time_for_64M_env_steps_s": (1 << 26) / (N * T / el)}), flush=True)
