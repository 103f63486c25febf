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


# ---- sdc-v4 (SDC_Full_Force_Env, sdc_force_env.py:7-118) --------------------------------------------------------
_cached_force = None


def force_reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sdc_gym", "envs", "sdc_force_env.py")) and reference_available()


def load_reference_force_env():
    """The unmodified ``sdc_force_env.py`` loaded by path.  It imports its base class relatively
    (``from .sdc_env import SDC_Full_Env``, ``sdc_force_env.py:4``), so the two files are mounted as submodules of a
    synthetic package instead of going through ``sdc_gym/__init__.py`` (gym registry, jax)."""
    global _cached_force
    if _cached_force is not None:
        return _cached_force
    base = load_reference_envs()
    pkg = types.ModuleType("_reference_envs_pkg")
    pkg.__path__ = []  # a package: relative imports resolve against sys.modules
    sys.modules["_reference_envs_pkg"] = pkg
    sys.modules["_reference_envs_pkg.sdc_env"] = base
    path = os.path.join(REFERENCE_ROOT, "sdc_gym", "envs", "sdc_force_env.py")
    spec = importlib.util.spec_from_file_location("_reference_envs_pkg.sdc_force_env", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_reference_envs_pkg.sdc_force_env"] = mod
    spec.loader.exec_module(mod)
    _cached_force = mod
    return mod


def make_reference_force_env(*, lam=None, **kwargs):
    """``SDC_Full_Force_Env`` with the ONE repair without which no non-diverging step of the reference completes:
    ``step`` calls ``self.reward_func(initial_residual, residual, done, niter)`` (``sdc_force_env.py:77-82``) although
    ``reward_func`` takes six positional arguments (``sdc_env.py:427-435``) - a TypeError.  The instance gets a
    ``reward_func`` that supplies ``None`` for the two omitted arguments (``scaled_action``, ``Pinv``: read only by the
    ``spectral_radius`` strategy, which therefore stays unusable); class, ``step`` and ``reset`` are the reference's."""
    mod = load_reference_force_env()
    env = mod.SDC_Full_Force_Env(**kwargs)
    orig = env.reward_func
    env.reward_func = lambda old, res, conv, steps, scaled_action=None, Pinv=None: orig(old, res, conv, steps,
                                                                                        scaled_action, Pinv)
    if lam is not None:
        env.reset()
        env.lam = complex(lam)
        env._compute_system_matrix()
        u = np.ones(env.M, dtype=np.complex128)
        residual = env._compute_residual(u)
        env.initial_residual = residual  # (reset(): _compute_initial_state sets it, sdc_env.py:306-314)
        env.state = (residual, np.zeros_like(u))
    return env
