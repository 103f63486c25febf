"""Experiment driver (GPU): time the sdc-v0 M=5 diag step kernel for each SDCGYM_TUNE variant.
Needs a library built with SDCGYM_TUNE_VARIANTS=1.  Output: one line per variant (ms per step, env-steps/s)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes

import numpy as np
import torch

import sdc_gym_b200

N = 1 << 20
mode = sys.argv[1] if len(sys.argv) > 1 else "uniform"
env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=5, dt=1.0, restol=1e-10, seed=0,
                        lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
gen = torch.Generator(device=env.device); gen.manual_seed(1)
if mode == "uniform":
    pool = [torch.rand((N, 5), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1 for _ in range(4)]
else:
    x = torch.tensor([0.2818591930905709, 0.2011358490453793, 0.06274536689514164, 0.11790265267514095,
                      0.1571629578515223], dtype=torch.float64, device=env.device)
    pool = [2 * (x[None] + (torch.rand((N, 5), dtype=torch.float64, device=env.device, generator=gen) - 0.5) * 0.06) - 1
            for _ in range(4)]
libc = ctypes.CDLL(None)
ref = None
for v in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,4,5,6,7,8,9".split(","))]:
    libc.setenv(b"SDCGYM_TUNE", str(v).encode(), 1)
    env.seed(0); env.episodes.zero_(); env.reset()
    for k in range(3):
        env.step_tensor(pool[k % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 20
    for k in range(K):
        out = env.step_tensor(pool[k % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    sig = (int(out["niter"].sum().item()), float(out["residual"].double().sum().item()))
    ref = ref or sig
    print(f"variant {v}: {ms:.4f} ms/step  {N / ms / 1e3:.1f} M env-steps/s  mean niter {sig[0] / N:.3f}  same_bits={sig == ref}", flush=True)
