"""Experiment driver (GPU): spectral-radius grid launches (for ncu / timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdc_gym_b200.loss import SpectralRadiusLoss
from sdc_gym_b200.precond import fixed_preconditioner
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
loss = SpectralRadiusLoss(5, 1.0, "diag")
x = np.diag(fixed_preconditioner("min", 5))
for _ in range(3):
    rho = loss.grid(G, G, [-100, 0], [-10, 0], x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    rho = loss.grid(G, G, [-100, 0], [-10, 0], x)
e1.record(); torch.cuda.synchronize()
print(f"grid {G}x{G}: {e0.elapsed_time(e1)/5:.3f} ms, {G*G/(e0.elapsed_time(e1)/5)/1e3:.1f} M matrices/s, mean rho {float(rho.mean()):.6f}")
