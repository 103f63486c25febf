"""Experiment driver (GPU): stage time stamps of the chunked host step (SDCGYM_PIPE_TRACE=1, csrc/hostpipe.cu) and the
split of a step's wall clock into the C call and the Python around it, 2^20 sdc-v0 envs."""
import os, sys, time
os.environ["SDCGYM_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdc_gym_b200
from sdc_gym_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
KW = dict(M=5, dt=1.0, restol=1e-10, seed=0, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
for chunks in (4, 8, 12, 16, 24):
    env = sdc_gym_b200.make("sdc-v0", num_envs=N, pipeline_chunks=chunks, **KW)
    env.reset()
    buf = env.pinned_action_buffer(0)
    buf[:] = np.random.default_rng(0).uniform(-1, 1, (N, 5))
    L = env._L
    orig = L.sdcgym_pipe_step_block
    tc = [0.0]

    def timed_call(*a):
        t = time.perf_counter()
        rc = orig(*a)
        tc[0] += time.perf_counter() - t
        return rc

    for _ in range(3):
        env.step(buf)
    sys.stderr.flush()
    os.environ["SDCGYM_PIPE_TRACE"] = "0"
    env._L = type("L", (), {"sdcgym_pipe_step_block": staticmethod(timed_call)})()
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        obs, rew, done, infos = env.step(buf)
    el = time.perf_counter() - t0
    print(f"chunks={chunks}: {el / K * 1e3:.3f} ms per step, of which the C call {tc[0] / K * 1e3:.3f} ms", flush=True)
    env._L = L
    del obs, rew, done, infos
    env.close()
