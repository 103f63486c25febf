"""Experiment driver (GPU): time the sdc-v1 M=5 diag step kernel for each SDCGYM_TUNE variant (occupancy sweep)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200

N = 1 << 22
libc = ctypes.CDLL(None)
for strat in ("iteration_only", "residual_change"):
    env = sdc_gym_b200.make("sdc-v1", num_envs=N, M=5, dt=1.0, restol=1e-10, seed=0, reward_strategy=strat,
                            lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    gen = torch.Generator(device=env.device); gen.manual_seed(1)
    pool = [torch.rand((N, 5), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1 for _ in range(2)]
    for v in range(0, 10):
        libc.setenv(b"SDCGYM_TUNE", str(v).encode(), 1)
        env.seed(0); env.episodes.zero_(); env.reset()
        for k in range(3):
            env.step_tensor(pool[k % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 40
        for k in range(K):
            out = env.step_tensor(pool[k % 2])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{strat} variant {v}: {ms:.4f} ms/step  {N / ms / 1e6:.2f} G env-steps/s  ({N * 437 / ms / 1e6:.0f} GB/s actual traffic est.)", flush=True)
    del env
