"""Experiment driver (GPU): a few certified-mode steps of the benchmark workload (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
M = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N = 1 << 20
env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0, sweep_mode="certified",
                        lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(1)
a = torch.rand((N, M), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1
for _ in range(3):
    out = env.step_tensor(a)
torch.cuda.synchronize()
print("mean niter", float(out["niter"].double().mean()), "fallback", env.fallback_stats())
