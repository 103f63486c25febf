"""Experiment driver (GPU): a few phased dense steps of one configuration (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
from sdc_gym_b200.precond import num_actions
M = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pt = sys.argv[2] if len(sys.argv) > 2 else "lower_tri"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, prec_type=pt, do_scale=False, dt=1.0, restol=1e-10, seed=0,
                        lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], phased=True)
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(1)
a = torch.rand((N, num_actions(M, pt)), dtype=torch.float64, device=env.device, generator=gen) * 0.3
for _ in range(3):
    out = env.step_tensor(a, want_terminal=False)
torch.cuda.synchronize()
print("mean niter", float(out["niter"].double().mean()), env.phase_count.cpu().numpy())
