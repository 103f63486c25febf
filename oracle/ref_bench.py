"""TEST / BASELINE INFRASTRUCTURE ONLY - times the reference's own CPU implementation of the hot path.

``DummyVecEnvLoop`` restates the loop the reference wraps its envs in (``utils/utils.py:284-294``: SB3 ``DummyVecEnv``
over ``TimeLimit(env)``; SB3 and gym are not installed - semantics restated from their documentation, as in
``oracle/sdc_port.py``) around env objects of either implementation:

* ``impl='reference'`` - the UNMODIFIED ``sdc_gym/envs/sdc_env.py`` loaded by ``oracle/ref_loader.py`` (live tree in the
  build container, the copy staged in ``oracle/_ref/`` on the GPU box);
* ``impl='port'``      - ``oracle/sdc_port.py`` (same numpy calls; used when no reference file is reachable).

Only ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and ``tests/`` import this module.
"""
from __future__ import annotations

import time

import numpy as np

from . import ref_loader, sdc_port

MAX_EPISODE_STEPS = {"sdc-v0": 1, "sdc-v1": 50}  # sdc_gym/__init__.py:3-13


def available_impl() -> str:
    return "reference" if ref_loader.reference_path() is not None else "port"


def _env_class(kind, impl):
    if impl == "reference":
        mod = ref_loader.load_reference_envs()
        return {"sdc-v0": mod.SDC_Full_Env, "sdc-v1": mod.SDC_Step_Env}[kind]
    return sdc_port.ENV_CLASSES[kind]


class DummyVecEnvLoop:
    def __init__(self, kind, num_envs, impl, seed=None, **kwargs):
        cls = _env_class(kind, impl)
        self.envs = [cls(seed=None if seed is None else seed + i, **kwargs) for i in range(num_envs)]
        self.num_envs, self.max_steps = num_envs, MAX_EPISODE_STEPS[kind]
        self._elapsed = [0] * num_envs

    def reset(self):
        self._elapsed = [0] * self.num_envs
        return np.stack([np.stack(e.reset()) for e in self.envs])

    def step(self, actions):
        obs, rews, dones, infos = [], [], [], []
        for i, (e, a) in enumerate(zip(self.envs, actions)):
            o, r, d, info = e.step(a)
            self._elapsed[i] += 1
            if self._elapsed[i] >= self.max_steps:  # gym TimeLimit
                info["TimeLimit.truncated"] = not d
                d = True
            if d:  # DummyVecEnv.step_wait
                info["terminal_observation"] = np.stack(o).copy()
                o = e.reset()
                self._elapsed[i] = 0
            obs.append(np.stack(o).copy())
            rews.append(r)
            dones.append(d)
            infos.append(info)
        return np.stack(obs), np.array(rews), np.array(dones), infos


def rollout_throughput(kind, seconds, *, impl=None, num_envs=8, M=5, seed=0, **kwargs):
    """Random-action rollout (BASELINE.json configs[0] / the CPU leg of configs[1]) for about ``seconds`` of wall
    clock.  Returns (env_steps, elapsed_s, sum_niter, impl)."""
    impl = impl or available_impl()
    kw = dict(M=M, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    kw.update(kwargs)
    vec = DummyVecEnvLoop(kind, num_envs, impl, seed=seed, **kw)
    rng = np.random.RandomState(seed + 12345)
    vec.reset()
    steps = sum_niter = 0
    t0 = time.perf_counter()
    while True:
        actions = rng.uniform(-1, 1, (num_envs, M))
        _, _, _, infos = vec.step(list(actions))
        steps += num_envs
        sum_niter += sum(i["niter"] for i in infos) if kind == "sdc-v0" else num_envs
        el = time.perf_counter() - t0
        if el >= seconds:
            return steps, el, sum_niter, impl


def spectral_radius_throughput(seconds, *, impl=None, M=5, seed=0):
    """The numpy spectral-radius path (``sdc_env.py:421-425`` ``_reward_spectral_radius`` after ``_compute_pinv``
    ``:193-201``) on random (lambda, MIN-diagonal) samples: matrices per second on one core."""
    impl = impl or available_impl()
    env = _env_class("sdc-v0", impl)(M=M, dt=1.0, restol=1e-10, prec="min", seed=seed,
                                     lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    env.reset()
    rng = np.random.RandomState(seed)
    n = 0
    acc = 0.0
    t0 = time.perf_counter()
    while True:
        for _ in range(256):
            env.lam = complex(rng.uniform(-100, 0), rng.uniform(-10, 0))
            pinv = env._compute_pinv(None)
            if impl == "reference":
                acc += env._reward_spectral_radius(None, pinv)
            else:  # the port keeps this inside reward_func: same statements as sdc_env.py:421-425
                acc += max(abs(np.linalg.eigvals(env.lam * env.dt * pinv.dot(env.Q - env._get_prec(None)))))
        n += 256
        el = time.perf_counter() - t0
        if el >= seconds:
            return n, el, acc / n, impl
