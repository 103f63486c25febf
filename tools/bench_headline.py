"""Experiment driver (GPU): the headline sdc-v0 step (M diag, uniform actions) timed alone with CUDA events.
Environment switches of the library (SDCGYM_STAGGER_NS, ...) are read once per process: run once per setting."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200

M = int(os.environ.get("M", 5))
N = int(os.environ.get("N", 1 << 20))
env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0,
                        lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(1)
acts = [torch.rand((N, M), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1 for _ in range(4)]
for k in range(5):
    env.step_tensor(acts[k % 4], want_terminal=False)
torch.cuda.synchronize()
best = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20):
        out = env.step_tensor(acts[k % 4], want_terminal=False)
    e1.record(); torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 20)
print(json.dumps({"M": M, "envs": N, "stagger_ns": os.environ.get("SDCGYM_STAGGER_NS", "0"), "ms": [round(b, 4) for b in best],
                  "mean_niter": round(float(out["niter"].double().mean()), 3)}), flush=True)
