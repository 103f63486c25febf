"""Experiment driver (GPU): cost split of the dense-inverse kernels - one sweep (max_iters=1) vs the full solve."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
from sdc_gym_b200.precond import num_actions
N = 1 << 21
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0)
for M in (5, 7, 9):
    for pt in ("diag", "lower_tri"):
        env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, prec_type=pt, do_scale=(pt == "diag"), **KW)
        A = num_actions(M, pt)
        a = torch.rand((N, A), dtype=torch.float64, device=dev, generator=gen) * (2 if pt == "diag" else 0.12) - (1 if pt == "diag" else 0)
        res = {}
        for mi in (1, 50):
            env._desc.max_iters = mi
            env.reset()
            for _ in range(2): env.step_tensor(a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): out = env.step_tensor(a)
            e1.record(); torch.cuda.synchronize()
            res[mi] = e0.elapsed_time(e1) / 5
        print(json.dumps({"M": M, "prec_type": pt, "envs": N, "ms_one_sweep": round(res[1], 3), "ms_full": round(res[50], 3),
                          "mean_niter_full": round(float(out["niter"].double().mean()), 2)}), flush=True)
        del env
