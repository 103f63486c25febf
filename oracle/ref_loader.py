"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference env module for validation / golden vectors.

``load_reference_envs()`` imports ``/root/reference/sdc_gym/envs/sdc_env.py`` by path (bypassing
``sdc_gym/__init__.py``, which needs gym's registry and jax) after installing three stub modules for the
imports that file makes and that are not installed in this image:

* ``gym`` / ``gym.spaces`` / ``gym.utils.seeding`` - only ``gym.Env``, ``spaces.Box`` and ``seeding.np_random``
  are touched (``sdc_env.py:4-6,15,89-110,119``);
* ``matplotlib.pyplot`` - imported, used only by ``plot_rewards`` (``sdc_env.py:7,465-496``);
* ``pySDC.implementations.collocation_classes.gauss_radau_right`` - ``CollGaussRadau_Right(M, 0, 1)``
  with ``.Qmat``, ``.delta_m``, ``.num_nodes`` (``sdc_env.py:11-12,53-54,186``), served by
  ``sdc_gym_b200.collocation`` so that reference and product share identical collocation bits.

The reference directory exists only in the build container, never on the GPU box.  Parity there rests on the
committed fixtures in ``tests/golden`` (made by ``tests/golden/make_golden.py`` through this loader) and the
travelling restatements ``oracle/sdc_port.py`` / ``oracle/sdc_exact.c``.  For TIMING the reference on the GPU box's
host cores (``bench.py --impl reference`` / ``cpu_baseline``), ``make -C oracle ref`` (run by
``__graft_entry__.build()`` in the build container) stages the unmodified ``sdc_env.py`` into the git-ignored
``oracle/_ref/``, which travels with the snapshot; ``reference_path()`` prefers the live tree and falls back to it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SDC_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


STAGED_REF = os.path.join(_REPO, "oracle", "_ref", "sdc_env.py")


def reference_available() -> bool:
    """the live reference tree (build container only) - what parity tests and golden generation need"""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sdc_gym", "envs", "sdc_env.py"))


def reference_path(allow_staged: bool = True):
    """path of the unmodified reference env module: the live tree, else the copy staged by ``make -C oracle ref``"""
    live = os.path.join(REFERENCE_ROOT, "sdc_gym", "envs", "sdc_env.py")
    if os.path.isfile(live):
        return live
    if allow_staged and os.path.isfile(STAGED_REF):
        return STAGED_REF
    return None


def _install_stubs():
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    from sdc_gym_b200.collocation import CollGaussRadauRight

    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # minimal gym.Env
            metadata = {}
            reward_range = (-float("inf"), float("inf"))

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

            def sample(self):
                raise NotImplementedError

        spaces = types.ModuleType("gym.spaces")
        spaces.Box = Box
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")

        def np_random(seed=None):
            # old-gym contract: (RandomState-like with .uniform, seed).  lambda is injected in all parity
            # runs, so the concrete stream is irrelevant ("parity unpinned" for gym's RNG).
            return np.random.RandomState(seed), seed

        seeding.np_random = np_random
        utils.seeding = seeding
        gym.Env, gym.spaces, gym.utils = Env, spaces, utils
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding})

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})

    name = "pySDC.implementations.collocation_classes.gauss_radau_right"
    if name not in sys.modules:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sys.modules.setdefault(".".join(parts[:i]), types.ModuleType(".".join(parts[:i])))
        sys.modules[name].CollGaussRadau_Right = CollGaussRadauRight


_cached = None


def load_reference_envs():
    """Return the reference module object (attributes ``SDC_Full_Env``, ``SDC_Step_Env``)."""
    global _cached
    if _cached is not None:
        return _cached
    path = reference_path()
    if path is None:
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT} nor staged at {STAGED_REF}")
    _install_stubs()
    spec = importlib.util.spec_from_file_location("_reference_sdc_env", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


def make_reference_env(kind: str, *, lam=None, **kwargs):
    """Construct a reference env ('sdc-v0' -> SDC_Full_Env, 'sdc-v1' -> SDC_Step_Env).

    If ``lam`` is given, the env is reset and then re-pointed at that lambda exactly the way the reference
    itself initialises an episode (``sdc_env.py:302-314``).
    """
    mod = load_reference_envs()
    cls = {"sdc-v0": mod.SDC_Full_Env, "sdc-v1": mod.SDC_Step_Env}[kind]
    env = cls(**kwargs)
    if lam is not None:
        env.reset()
        inject_lambda(env, lam)
    return env


def inject_lambda(env, lam):
    """Overwrite the freshly reset episode of ``env`` with a chosen lambda (same statements as reset())."""
    env.lam = complex(lam)
    env._compute_system_matrix()
    u = np.ones(env.M, dtype=np.complex128)
    residual = env._compute_residual(u)
    env.initial_residual = residual
    env.state = (u, residual)
    if env.collect_states:
        env.old_states[:, 0] = np.concatenate(env.state)
        env.old_states[:, 1:] = 0
    return env.state
