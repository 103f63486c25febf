"""Experiment driver (GPU): sdc-v0 / sdc-v1 throughput of the dense-Q_delta kernels at M = 6..9.
SDCGYM_NO_TEAM=1 selects the per-thread kernels instead of the lane-team kernels (A/B)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
from sdc_gym_b200.precond import num_actions
N = 1 << 21
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0)
cases = [(M, "lower_tri", None, "sdc-v0") for M in (6, 7, 8, 9)] + [(7, "diag", "LU", "sdc-v0"), (9, "diag", "LU", "sdc-v0"),
                                                                     (7, "lower_tri", None, "sdc-v1")]
for M, pt, prec, name in cases:
    env = sdc_gym_b200.make(name, num_envs=N, M=M, prec=prec, prec_type=pt, do_scale=False, **KW)
    env.reset()
    a = None if prec else torch.rand((N, num_actions(M, pt)), dtype=torch.float64, device=dev, generator=gen) * 0.3
    for _ in range(3): env.step_tensor(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = env.step_tensor(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"env": name, "M": M, "prec_type": pt, "prec": prec, "team": os.environ.get("SDCGYM_NO_TEAM") is None,
                      "ms": round(ms, 3), "Menv_steps_per_s": round(N / ms / 1e3, 1),
                      "mean_niter": round(float(out["niter"].double().mean()), 2)}), flush=True)
    del env
