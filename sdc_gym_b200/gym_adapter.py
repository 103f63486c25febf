"""``gym.Env``-shaped adapter over a one-env ``SDCVecEnv`` - what ``gym.make('sdc-v0' | 'sdc-v1', **kwargs)`` returns
after ``sdc_gym_b200.register_gym()`` (reference registration: ``sdc_gym/__init__.py:3-13``; env protocol:
``sdc_gym/envs/sdc_env.py:209-273,316-332,507-572``).

The adapter exists for the reference's scripts that build envs through gym's registry; it steps ONE env per call and is
therefore latency bound (~50 us per step).  Throughput comes from ``sdc_gym_b200.make(envname, num_envs=N)``.
``gym`` itself is optional: without it the class still works as a plain object with ``reset`` / ``step``.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    import gym as _gym

    _Base = _gym.Env
except Exception:  # pragma: no cover
    try:
        import gymnasium as _gym

        _Base = _gym.Env
    except Exception:
        _Base = object


class SingleEnv(_Base):
    """One reference env: ``reset() -> (u, r)``, ``step(action) -> ((u, r), reward, done, info)`` (old-gym 4-tuple).
    No auto-reset (gym's ``TimeLimit`` / ``DummyVecEnv`` add theirs on top, as they do for the reference env)."""

    metadata = {"render.modes": []}

    def __init__(self, envname="sdc-v0", **kwargs):
        from .vec_env import SDCVecEnv

        kwargs.pop("num_envs", None)
        self._vec = SDCVecEnv(envname, num_envs=1, autoreset=False, **kwargs)
        self.observation_space = self._vec.observation_space
        self.action_space = self._vec.action_space

    # attributes the reference's scripts read (rl_playground.py:147-156, dp_playground.py:742-749)
    prec = property(lambda self: self._vec.prec)
    restol = property(lambda self: self._vec.restol)
    M = property(lambda self: self._vec.M)
    dt = property(lambda self: self._vec.dt)
    Q = property(lambda self: self._vec.Q)
    lam = property(lambda self: self._vec.envs[0].lam)
    state = property(lambda self: self._vec.envs[0].state)
    niter = property(lambda self: self._vec.envs[0].niter)
    initial_residual = property(lambda self: self._vec.envs[0].initial_residual)
    num_episodes = property(lambda self: self._vec.envs[0].num_episodes)

    def set_num_episodes(self, n):
        self._vec.set_num_episodes(n)

    def seed(self, seed=None):
        self._vec.seed(seed)
        return [seed]

    def _obs(self, obs):
        if self._vec.collect_states:
            return obs[0]
        return (obs[0, 0], obs[0, 1])

    def reset(self, **kwargs):
        return self._obs(self._vec.reset(**kwargs))

    def step(self, action):
        a = None if self._vec.prec is not None else np.asarray(action).reshape(1, -1)
        obs, rew, done, infos = self._vec.step(a)
        # the env's own info dict (sdc_env.py:265-269); 'terminal_observation' / 'TimeLimit.truncated' are added by
        # DummyVecEnv / gym's TimeLimit wrapper on top, as for the reference env
        info = {"residual": infos.residual[0], "niter": int(infos.niter[0]), "lam": complex(infos.lam[0])}
        return self._obs(obs), float(rew[0]), bool(done[0]), info

    def close(self):
        self._vec.close()


class SingleForceEnv(_Base):
    """One ``SDC_Full_Force_Env`` (``sdc_force_env.py:7-118``): ``reset() -> (residual, zeros)``,
    ``step(action) -> ((residual, diagonal), reward, done, info)`` with ``info['ntries']``."""

    metadata = {"render.modes": []}

    def __init__(self, envname="sdc-v4", **kwargs):
        from .force_env import SDCForceVecEnv

        kwargs.pop("num_envs", None)
        self._vec = SDCForceVecEnv(envname, num_envs=1, autoreset=False, **kwargs)
        self.observation_space = self._vec.observation_space
        self.action_space = self._vec.action_space
        self.max_tries = self._vec.max_tries

    prec = property(lambda self: self._vec.prec)
    restol = property(lambda self: self._vec.restol)
    M = property(lambda self: self._vec.M)
    ntries = property(lambda self: int(self._vec.ntries[0]))

    def seed(self, seed=None):
        self._vec.seed(seed)
        return [seed]

    def reset(self, **kwargs):
        obs = self._vec.reset(**kwargs)
        return (obs[0, 0], obs[0, 1])

    def step(self, action):
        obs, rew, done, infos = self._vec.step(np.asarray(action, dtype=np.float64).reshape(1, -1))
        return (obs[0, 0], obs[0, 1]), float(rew[0]), bool(done[0]), infos[0]

    def close(self):
        self._vec.close()


def make_single(envname="sdc-v0", **kwargs):
    if envname == "sdc-v4":
        return SingleForceEnv(envname, **kwargs)
    return SingleEnv(envname, **kwargs)
