"""Experiment driver (GPU): wall-clock latency / throughput of `SDCVecEnv.step(numpy)` (host arrays in and out, default
settings) from 1 env to 2^20 envs, for the library's own chunk choice and for forced chunk counts - the data behind
`auto_chunks` in csrc/hostpipe.cu.  One JSON line per (env id, batch size, chunks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdc_gym_b200

KW = dict(M=5, dt=1.0, restol=1e-10, seed=0, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
for name in ("sdc-v0", "sdc-v1"):
    for n in (1, 8, 64, 1024, 4096, 16384, 32768, 65536, 131072, 262144, 524288, 1 << 20):
        variants = [0] if n < 4096 else [0, 1, 2, 4, 8]
        for chunks in variants:
            env = sdc_gym_b200.make(name, num_envs=n, pipeline_chunks=chunks, **KW)
            env.reset()
            buf = env.pinned_action_buffer(0)
            buf[:] = np.random.default_rng(0).uniform(-1, 1, (n, 5))
            K = 1000 if n <= 1024 else (200 if n <= 65536 else 30)
            for _ in range(max(5, K // 10)):
                env.step(buf)
            t0 = time.perf_counter()
            for _ in range(K):
                obs, rew, done, infos = env.step(buf)
            dt = (time.perf_counter() - t0) / K
            print(json.dumps({"env": name, "num_envs": n, "chunks": chunks or "auto", "us_per_step": round(dt * 1e6, 1),
                              "env_steps_per_s": round(n / dt), "host_blocks": len(env._host["sets"]),
                              "copies": env.host_set_copies}), flush=True)
            del obs, rew, done, infos
            env.close()
            del env
