"""Experiment driver (GPU): wall-clock latency of SDCVecEnv.step(numpy) for small batches (the reference's own regime:
num_envs = 1..64 on the CPU), host arrays in and out."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdc_gym_b200
for name in ("sdc-v1", "sdc-v0"):
    for n in (1, 8, 64, 1024, 4096, 16384, 65536, 131072):
        env = sdc_gym_b200.make(name, num_envs=n, M=5, dt=1.0, restol=1e-10, seed=0,
                                lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
        env.reset()
        a = np.random.default_rng(0).uniform(-1, 1, (n, 5))
        for _ in range(50): env.step(a)
        t0 = time.perf_counter()
        K = 2000 if n <= 1024 else 300
        for _ in range(K): obs, rew, done, infos = env.step(a)
        dt = (time.perf_counter() - t0) / K
        print(json.dumps({"env": name, "num_envs": n, "us_per_step": round(dt * 1e6, 1), "env_steps_per_s": round(n / dt)}), flush=True)
# the reference's training regime: DummyVecEnv of 8 envs wrapped in VecNormalize (utils/utils.py:295-312), numpy in/out
for n in (8, 64, 1024):
    env = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=n, M=5, dt=1.0, restol=1e-10, seed=0,
                                                      lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0]))
    env.reset()
    a = np.random.default_rng(0).uniform(-1, 1, (n, 5))
    for _ in range(50): env.step(a)
    t0 = time.perf_counter()
    K = 1000
    for _ in range(K): obs, rew, done, infos = env.step(a)
    dt = (time.perf_counter() - t0) / K
    print(json.dumps({"env": "VecNormalize(sdc-v1)", "num_envs": n, "us_per_step": round(dt * 1e6, 1),
                      "env_steps_per_s": round(n / dt)}), flush=True)
