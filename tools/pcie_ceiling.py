"""What the box's host<->device links allow (GPU tool): N ranks move the bytes of one `SDCVecEnv.step` (2^20 envs,
M = 5: 42 MB of actions up, 123 MB of results down) with bare, concurrent cudaMemcpyAsync from/to page-locked memory -
no kernels - and report GB/s per GPU, aggregate, and the env-steps/s those copies alone would allow.  `bench.py`
measures the same ceiling inside every run (`e2e.pcie_ceiling`); this tool adds the per-direction numbers.

    python tools/pcie_ceiling.py                                                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    dist.init_process_group("nccl", device_id=dev)
N, M = args.envs, 5
h2d_bytes, d2h_bytes = N * M * 8, N * (M * 16 + 8 + 1 + 4 + 8 + 16)
up_h = torch.zeros(h2d_bytes, dtype=torch.uint8, pin_memory=True); up_d = torch.zeros(h2d_bytes, dtype=torch.uint8, device=dev)
dn_h = torch.zeros(d2h_bytes, dtype=torch.uint8, pin_memory=True); dn_d = torch.zeros(d2h_bytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def run(up, down):
    def once():
        if up:
            with torch.cuda.stream(s1):
                up_d.copy_(up_h, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                dn_h.copy_(dn_d, non_blocking=True)
        s1.synchronize(); s2.synchronize()
    for _ in range(3):
        once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        once()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / args.iters


res = {}
for name, up, down in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
    s = run(up, down)
    res[name] = {"ms": s * 1e3, "h2d_GBps_per_gpu": h2d_bytes / s / 1e9 if up else None,
                 "d2h_GBps_per_gpu": d2h_bytes / s / 1e9 if down else None,
                 "aggregate_GBps": world * ((h2d_bytes if up else 0) + (d2h_bytes if down else 0)) / s / 1e9}
if rank == 0:
    print(json.dumps({"tool": "pcie_ceiling", "n_gpus": world, "envs_per_gpu": N, "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes,
                      **res, "env_steps_per_s_ceiling": world * N / (res["both"]["ms"] * 1e-3)}), flush=True)
if world > 1:
    dist.destroy_process_group()
