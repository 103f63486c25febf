"""Experiment driver (GPU): where the end-to-end step time goes (PCIe copies vs kernels vs host overhead)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdc_gym_b200

N, M = 1 << 20, 5
dev = torch.device("cuda", 0)
def timeit(f, n=10):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3
a_h = torch.zeros((N, M), dtype=torch.float64, pin_memory=True); a_d = torch.zeros((N, M), dtype=torch.float64, device=dev)
o_h = torch.zeros((N, 2, M, 2), dtype=torch.float64, pin_memory=True); o_d = torch.zeros((N, 2, M, 2), dtype=torch.float64, device=dev)
print(f"H2D 42 MB pinned: {timeit(lambda: a_d.copy_(a_h, non_blocking=True)):.3f} ms")
print(f"D2H 168 MB pinned: {timeit(lambda: o_h.copy_(o_d, non_blocking=True)):.3f} ms")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): a_d.copy_(a_h, non_blocking=True)
    with torch.cuda.stream(s2): o_h.copy_(o_d, non_blocking=True)
print(f"H2D + D2H concurrent: {timeit(both):.3f} ms")
x = np.zeros((N, M)); y = np.ones((N, M))
t = time.perf_counter(); [np.copyto(x, y) for _ in range(5)]; print(f"host memcpy 42 MB: {(time.perf_counter()-t)/5*1e3:.3f} ms")
for chunks in (1, 2, 4, 8, 16):
    env = sdc_gym_b200.make("sdc-v0", num_envs=N, M=M, dt=1.0, restol=1e-10, seed=0, lambda_real_interval=[-100, 0],
                            lambda_imag_interval=[-10, 0], reuse_buffers=True, pipeline_chunks=chunks)
    env.reset()
    buf = env.pinned_action_buffer(); buf[:] = np.random.default_rng(0).uniform(-1, 1, (N, M))
    ms = timeit(lambda: env.step(buf), 8)
    print(f"chunks={chunks:2d}: step {ms:.3f} ms  -> {N/ms/1e3:.1f} M env-steps/s")
    del env
