"""Experiment driver (GPU): a few normalised sdc-v1 steps (device VecNormalize) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
N, M = 1 << 20, 5
env = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0,
                                                  reward_iteration_only=False, lambda_real_interval=[-100, 0],
                                                  lambda_imag_interval=[-10, 0]), norm_obs=True, norm_reward=True)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
a = torch.rand((N, M), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
for _ in range(4):
    out = env.step_tensor(a)
torch.cuda.synchronize()
print("ok", float(out["reward"].mean()))
