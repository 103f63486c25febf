"""BASELINE config 5 at scale: sdc-v1 rollout collection with the device VecNormalize (statistics synchronised over
NCCL every step: one < 1 kB all-reduce per statistic), RolloutBuffer and GAE, envs sharded over the ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_rollout_dist.py

Default: 2^20 envs per GPU x 8 steps = 64M env-steps per rollout on 8 GPUs.  Rank 0 prints one JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import sdc_gym_b200
from sdc_gym_b200.rollout import collect_rollouts

ap = argparse.ArgumentParser()
ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--rollouts", type=int, default=5)
ap.add_argument("--sync", default="peer", choices=["peer", "nccl"], help="peer: all-reduce of the moment sums inside the statistics kernel over NVLink peer memory; nccl: accumulate -> ncclAllReduce -> merge")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    dist.init_process_group("nccl", device_id=dev)
N, T, M = args.envs_per_gpu, args.steps, 5
env = sdc_gym_b200.VecNormalize(
    sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0, env_offset=rank * N,
                      reward_iteration_only=False, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
                      device=dev), norm_obs=True, norm_reward=True, sync=True if args.sync == "peer" else "nccl")
env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(1 + rank)


def policy(obs_planes):
    a = torch.empty((N, M), dtype=torch.float64, device=dev).uniform_(-1.0, 1.0, generator=gen)
    return a, obs_planes[0], None


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


buf = collect_rollouts(env, policy, T)
buf = collect_rollouts(env, policy, T, buffer=buf)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.rollouts):
    buf = collect_rollouts(env, policy, T, buffer=buf)
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / args.rollouts], dtype=torch.float64, device=dev)
mean0 = env.obs_rms.mean.clone()
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = mean0.clone(); dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(ref, mean0))], device=dev); dist.all_reduce(same, op=dist.ReduceOp.MIN)
else:
    same = torch.ones(1)
if rank == 0:
    total = world * N * T
    print(json.dumps({"config": "sdc-v1 rollout collection, device VecNormalize synchronised over all ranks every step, RolloutBuffer + GAE",
                      "n_gpus": world, "normaliser_sync": args.sync if world > 1 else "single rank", "envs_per_gpu": N, "n_steps": T, "env_steps_per_rollout": total,
                      "ms_per_rollout": float(ms.item()), "env_steps_per_s": total / (float(ms.item()) * 1e-3),
                      "normaliser_identical_on_all_ranks": bool(same.item() == 1.0)}), flush=True)
if world > 1:
    dist.destroy_process_group()
