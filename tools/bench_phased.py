"""Experiment driver (GPU): phased dense full solve (step_one PHASE) against the single launch, BASELINE config 3
workload (Q_delta entries ~ U[0, 0.3], 2^22 envs).  `SDCGYM_PHASE_STOPS=a,b,...` picks the hand-over sweep counts
(read once per process: run once per schedule).  Prints one JSON line per case."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
from sdc_gym_b200.precond import num_actions

N = int(os.environ.get("N", 1 << 22))
cases = [(int(m), pt) for m in os.environ.get("MS", "3,5,7").split(",") for pt in ("lower_tri", "strictly_lower_tri")]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0)
for M, pt in cases:
    a = torch.rand((N, num_actions(M, pt)), dtype=torch.float64, device=dev, generator=gen) * 0.3
    rec = {"M": M, "prec_type": pt, "envs": N, "stops": os.environ.get("SDCGYM_PHASE_STOPS", "default")}
    for phased in (False, True):
        env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, prec_type=pt, do_scale=False, phased=phased, **KW)
        env.reset()
        for _ in range(3): env.step_tensor(a, want_terminal=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): out = env.step_tensor(a, want_terminal=False)
        e1.record(); torch.cuda.synchronize()
        rec["phased_ms" if phased else "single_ms"] = round(e0.elapsed_time(e1) / 8, 3)
        if phased:
            rec["suspended"] = [int(c) for c in env.phase_count.cpu().numpy()[:6]]
        rec["mean_niter"] = round(float(out["niter"].double().mean()), 2)
        del env
    rec["speedup"] = round(rec["single_ms"] / rec["phased_ms"], 3)
    print(json.dumps(rec), flush=True)
