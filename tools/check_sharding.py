"""Multi-GPU correctness check (run under torchrun, one rank per GPU):
  1. every rank's shard of a global batch reproduces, bit for bit, its slice of the same batch stepped on one GPU;
  2. VecNormalize statistics synchronised over NCCL equal the single-GPU statistics to reduction-order accuracy;
  3. RolloutStats.reduce() equals the single-GPU totals;
  4. the spectral-radius grid sharded by rows (SpectralRadiusLoss.grid_mean, one scalar all-reduced) has the mean of
     the full grid evaluated on one GPU.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharding.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import sdc_gym_b200
from sdc_gym_b200 import dist as sdist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, ws = sdist.world()
NG = 200_003  # deliberately not divisible
KW = dict(M=5, dt=1.0, restol=1e-10, seed=7, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          reward_iteration_only=False)
off, cnt = sdist.shard_range(NG, rank, ws)
gen = torch.Generator(device="cuda"); gen.manual_seed(3)
acts = [(torch.rand((NG, 5), dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 0.3 - 0.65 for _ in range(6)]
ok = True
for kind in ("sdc-v0", "sdc-v1"):
    full = sdc_gym_b200.make(kind, num_envs=NG, **KW)                 # the whole batch on this GPU
    shard = sdist.make_sharded(kind, NG, **KW)                         # this rank's slice
    assert shard.num_envs == cnt
    vn_full = sdc_gym_b200.VecNormalize(full, sync=False)
    vn_shard = sdc_gym_b200.VecNormalize(shard, sync=True)
    vn_full.reset(); vn_shard.reset()
    stats = sdist.RolloutStats("cuda"); tot = sdist.RolloutStats("cuda")
    for a in acts:
        of = vn_full.step_tensor(a)
        os_ = vn_shard.step_tensor(a[off:off + cnt])
        stats.update(os_); tot.update(of)
        for k in ("raw_reward", "niter", "residual", "flags", "lam"):
            ok &= bool(torch.equal(of[k][off:off + cnt], os_[k]))
        ok &= bool(torch.equal(full.S[:, off:off + cnt], shard.S[:, :cnt]))
    red = stats.reduce()
    tot_local = dict(zip(tot.FIELDS, tot.acc.cpu().numpy()))
    for k in red:
        ok &= abs(red[k] - tot_local[k]) <= 1e-9 * max(1.0, abs(tot_local[k]))
    dm = float((vn_full.obs_rms.mean - vn_shard.obs_rms.mean).abs().max())
    dv = float(((vn_full.obs_rms.var - vn_shard.obs_rms.var).abs() / vn_full.obs_rms.var).max())
    dr = float(((vn_full.ret_rms.var - vn_shard.ret_rms.var).abs() / vn_full.ret_rms.var).max())
    ok &= dm < 1e-11 and dv < 1e-11 and dr < 1e-11 and abs(vn_full.obs_rms.count - vn_shard.obs_rms.count) < 1e-6
    if rank == 0:
        print(json.dumps({"kind": kind, "world": ws, "global_envs": NG, "steps": len(acts), "bit_equal_shards": bool(ok),
                          "obs_mean_maxabs_diff": dm, "obs_var_maxrel_diff": dv, "ret_var_rel_diff": dr,
                          "rollout": red}), flush=True)
# ---- 4. rho grid sharded by rows ----
from sdc_gym_b200.loss import SpectralRadiusLoss
from sdc_gym_b200.precond import fixed_preconditioner
loss = SpectralRadiusLoss(5, 1.0, "diag")
x = np.diag(fixed_preconditioner("min", 5))
G_RE, G_IM = 1001, 257  # rows not divisible by the world size
full_grid = loss.grid(G_RE, G_IM, [-100, 0], [-10, 0], x)
lo, cnt_rows = sdist.shard_range(G_RE, rank, ws)
mine = loss.grid(G_RE, G_IM, [-100, 0], [-10, 0], x, rows=(lo, lo + cnt_rows))
ok &= bool(torch.equal(mine, full_grid[lo:lo + cnt_rows]))
m_sharded = float(loss.grid_mean(G_RE, G_IM, [-100, 0], [-10, 0], x))
m_full = float(loss.mean(full_grid.reshape(-1)))
ok &= abs(m_sharded - m_full) <= 1e-13 * abs(m_full)
if rank == 0:
    print(json.dumps({"kind": "rho grid", "world": ws, "grid": [G_RE, G_IM], "rows_bit_equal": bool(ok),
                      "mean_sharded": m_sharded, "mean_single_gpu": m_full}), flush=True)
flag = torch.tensor([1.0 if bool(ok) else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL RANKS OK" if flag.item() == 1.0 else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
