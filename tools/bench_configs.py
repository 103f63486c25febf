"""Secondary measurements (GPU): the BASELINE.json configs that are not the headline bench line.

  config 2: M sweep 3/5/7/9 x {diag, lower_tri, strictly_lower_tri} sdc-v0, N envs per GPU
  config 3: spectral-radius loss over a grid_n x grid_n lambda grid, M=5 MIN diag (+ learned complex diag batch)
  config 4: sdc-v1 rollout collection with device VecNormalize(norm_obs), random policy
  plus sdc-v1 raw step throughput (HBM-bound kernel) with its achieved GB/s.

Prints one JSON object per measurement; `python tools/bench_configs.py [--envs N] [--grid G] > profiles/...`."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdc_gym_b200
from sdc_gym_b200.precond import fixed_preconditioner, num_actions

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 22)
ap.add_argument("--grid", type=int, default=4096)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda", 0)
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0)
HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


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


def actions_for(M, prec_type, N, gen):
    A = num_actions(M, prec_type)
    if prec_type == "diag":
        return torch.rand((N, A), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    return torch.rand((N, A), dtype=torch.float64, device=dev, generator=gen) * 0.3


gen = torch.Generator(device=dev); gen.manual_seed(1)
# ---- config 2 ----
for M in (3, 5, 7, 9):
    for pt in ("diag", "lower_tri", "strictly_lower_tri"):
        N = args.envs
        env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, prec_type=pt, do_scale=(pt == "diag"), **KW)
        env.reset()
        acts = [actions_for(M, pt, N, gen) for _ in range(2)]
        k = [0]
        def f():
            k[0] += 1
            return env.step_tensor(acts[k[0] % 2])
        ms = timed(f, args.steps)
        out = env.step_tensor(acts[0])
        flags = out["flags"]
        print(json.dumps({"config": "sdc-v0 M sweep", "M": M, "prec_type": pt, "envs": N, "ms_per_step": ms,
                          "env_steps_per_s": N / ms * 1e3, "mean_niter": float(out["niter"].double().mean()),
                          "converged_frac": float((flags & 2).ne(0).double().mean()),
                          "diverged_frac": float((flags & 4).ne(0).double().mean())}), flush=True)
        del env, acts
        torch.cuda.empty_cache()
# ---- fixed preconditioners ----
for prec in ("LU", "min"):
    N = args.envs
    env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=5, prec=prec, **KW)
    env.reset()
    ms = timed(lambda: env.step_tensor(None), args.steps)
    out = env.step_tensor(None)
    print(json.dumps({"config": "sdc-v0 fixed prec", "M": 5, "prec": prec, "envs": N, "ms_per_step": ms,
                      "env_steps_per_s": N / ms * 1e3, "mean_niter": float(out["niter"].double().mean()),
                      "converged_frac": float((out["flags"] & 2).ne(0).double().mean())}), flush=True)
    del env
    torch.cuda.empty_cache()
# ---- sdc-v1 raw step (HBM-bound) ----
for M in (5,):
    N = args.envs
    env = sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, reward_iteration_only=False, **KW)
    env.reset()
    acts = [actions_for(M, "diag", N, gen) for _ in range(2)]
    k = [0]
    def f1():
        k[0] += 1
        return env.step_tensor(acts[k[0] % 2])
    ms = timed(f1, 50)
    bytes_alg = 8 * M + 64 * M + 49
    print(json.dumps({"config": "sdc-v1 step", "M": M, "envs": N, "ms_per_step": ms, "env_steps_per_s": N / ms * 1e3,
                      "algorithmic_GBps": N * bytes_alg / ms / 1e6, "hbm_frac": N * bytes_alg / ms / 1e6 / HBM}), flush=True)
    # ---- config 4: rollout collection with device VecNormalize ----
    vn = sdc_gym_b200.VecNormalize(env, norm_obs=True, norm_reward=True)
    vn.reset()
    def f2():
        k[0] += 1
        return vn.step_tensor(acts[k[0] % 2])
    ms = timed(f2, 50)
    print(json.dumps({"config": "sdc-v1 rollout + device VecNormalize(norm_obs, norm_reward)", "M": M, "envs": N,
                      "ms_per_step": ms, "env_steps_per_s": N / ms * 1e3,
                      "time_for_64M_env_steps_s": (1 << 26) / (N / ms * 1e3)}), flush=True)
    # full rollout collection: random policy, observations/actions/rewards stored in a device RolloutBuffer, GAE
    from sdc_gym_b200.rollout import RolloutBuffer, collect_rollouts
    Nr, T = 1 << 20, 16
    env_r = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=Nr, M=M, reward_iteration_only=False, **KW))
    env_r.reset()
    def policy(obs_planes):
        a = torch.rand((Nr, M), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        return a, obs_planes[0], None
    buf = collect_rollouts(env_r, policy, T)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        buf = collect_rollouts(env_r, policy, T, buffer=buf)
    torch.cuda.synchronize()
    el = (time.perf_counter() - t0) / 4
    print(json.dumps({"config": "sdc-v1 rollout collection (VecNormalize + RolloutBuffer + GAE, random policy)", "M": M,
                      "envs": Nr, "n_steps": T, "s_per_rollout": el, "env_steps_per_s": Nr * T / el,
                      "time_for_64M_env_steps_s": (1 << 26) / (Nr * T / el)}), flush=True)
    del env, vn, env_r, buf
    torch.cuda.empty_cache()
# ---- config 3: spectral radius grid ----
from sdc_gym_b200.loss import SpectralRadiusLoss
G = args.grid
loss = SpectralRadiusLoss(5, 1.0, "diag")
x = np.diag(fixed_preconditioner("min", 5))
ms = timed(lambda: loss.grid(G, G, [-100, 0], [-10, 0], x), 5, warm=2)
rho = loss.grid(G, G, [-100, 0], [-10, 0], x)
print(json.dumps({"config": "spectral radius grid", "M": 5, "grid": [G, G], "prec": "MIN diag", "ms": ms,
                  "matrices_per_s": G * G / ms * 1e3, "mean_rho": float(loss.mean(rho.reshape(-1))),
                  "max_rho": float(rho.max()), "nan": int(torch.isnan(rho).sum())}), flush=True)
B = 1 << 22
lam = torch.complex(torch.rand(B, dtype=torch.float64, device=dev, generator=gen) * -100,
                    torch.rand(B, dtype=torch.float64, device=dev, generator=gen) * -10)
outp = torch.complex(torch.rand((B, 5), dtype=torch.float64, device=dev, generator=gen) * 0.4,
                     (torch.rand((B, 5), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 0.1)
ms = timed(lambda: loss.spectral_radii(lam, outp), 5, warm=2)
print(json.dumps({"config": "spectral radius batch (learned complex diag)", "M": 5, "batch": B, "ms": ms,
                  "matrices_per_s": B / ms * 1e3}), flush=True)
