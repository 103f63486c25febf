"""Experiment driver (GPU): headline step with C in tensor memory (SDCGYM_TMEM=1) against a recorded run of the shipped
kernel: `python tools/check_tmem.py record f.pt` (SDCGYM_TMEM unset), then `SDCGYM_TMEM=1 python tools/check_tmem.py check f.pt`."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200
mode, path = sys.argv[1], sys.argv[2]
M = int(os.environ.get("M", 5)); N = int(os.environ.get("N", 1 << 20))
env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0,
                        lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(1)
acts = [torch.rand((N, M), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1 for _ in range(3)]
outs = []
for a in acts:
    o = env.step_tensor(a)
    outs.append({k: v.clone().cpu() for k, v in o.items()})
    outs[-1]["S"] = env.S.clone().cpu()
torch.cuda.synchronize()
if mode == "record":
    torch.save(outs, path); print("recorded")
else:
    ref = torch.load(path)
    for k, (a, b) in enumerate(zip(ref, outs)):
        for name in a:
            x, y = a[name], b[name]
            if x.is_floating_point(): x, y = x.view(torch.int64), y.view(torch.int64)
            elif x.is_complex(): x, y = torch.view_as_real(x).view(torch.int64), torch.view_as_real(y).view(torch.int64)
            assert torch.equal(x, y), (k, name)
    print("bit-identical to the recorded run over", len(outs), "steps")
