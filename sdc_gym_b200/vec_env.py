"""``SDCVecEnv`` - the batched, device-resident replacement for
``DummyVecEnv([lambda: gym.make('sdc-v0'|'sdc-v1', seed=seed+i, **kwargs) for i in range(num_envs)])``
(reference ``utils/utils.py:284-294`` over ``sdc_gym/envs/sdc_env.py``).

All env state lives in HBM as planes (see ``include/sdcgym.h``); ``reset``/``step`` enqueue hand-written
sm_100a kernels through the C ABI of ``libsdcgym.so``.  torch is used only for device memory, streams and
pinned host buffers.  There is no CPU implementation behind this class.

API kept from the reference's callers (``rl_playground.py``, ``dp_playground.py``):
``reset() -> obs``, ``step(actions) -> (obs, rewards, dones, infos)`` (old-gym 4-tuple, DummyVecEnv
auto-reset with ``info['terminal_observation']``), ``step_async/step_wait``, ``seed``, ``num_envs``,
``observation_space``, ``action_space``, ``envs[i].{prec,lam,restol,M,state,initial_residual,num_episodes,
set_num_episodes}``, ``get_attr/set_attr/env_method``, ``close``.  Constructor keyword names are the
reference's (``sdc_env.py:27-46``) plus ``prec_type`` (``dp_playground.py:194-207`` layouts).
"""
from __future__ import annotations

import ctypes
import os
import sys
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .collocation import collocation_matrix
from .precond import fixed_preconditioner, num_actions
from .spaces import Box

MAX_ITERS = 50  # SDC_Full_Env.max_iters, sdc_env.py:25
PHASED_MAX_M = 7  # the phased dense solve covers the one-env-per-thread kernels (M >= 8: lane-team kernel)
PHASED_MIN_ENVS = 16384  # smallest batch that allocates the work buffers by default (phased=True: any batch)
PHASED_TRIAL = 3  # phased=None: launches of each sequence per measurement (the first one untimed)
PHASED_RETUNE = 512  # ... and device steps between two measurements
MAX_EPISODE_STEPS = {"sdc-v0": 1, "sdc-v1": 50}  # sdc_gym/__init__.py:3-13


def _torch():
    import torch

    return torch


def detect_blas_variant() -> int:
    """Which OpenBLAS core the host numpy dispatches to (decides the scalar-tail rounding)."""
    try:
        from threadpoolctl import threadpool_info

        for info in threadpool_info():
            if info.get("internal_api") == "openblas":
                arch = str(info.get("architecture", "")).lower()
                if arch in ("haswell", "zen", "sandybridge", "nehalem", "prescott", "core2"):
                    return _lib.BLAS_HASWELL
                return _lib.BLAS_SKYLAKEX
    except Exception:  # pragma: no cover
        pass
    return _lib.BLAS_SKYLAKEX


class LazyInfos(Sequence):
    """Sequence of per-env info dicts (``sdc_env.py:265-269,564-568``) materialised on access.

    Array views: ``.niter``, ``.residual``, ``.lam``, ``.done``; ``terminal_observation`` and
    ``TimeLimit.truncated`` appear for finished envs exactly as DummyVecEnv / TimeLimit add them.
    """

    def __init__(self, niter, residual, lam, done, truncated_key, terminal_fetch, info_fetch=None):
        self._niter, self._residual, self._lam, self.done = niter, residual, lam, done
        self._truncated_key = truncated_key
        self._terminal_fetch = terminal_fetch
        self._terminal = None
        # large pipelined batches leave niter / residual / lam (28 of 117 bytes per env) on the device until somebody
        # reads them: `info_fetch` copies them into the (already allocated) arrays above
        self._info_fetch = info_fetch

    def _info(self):
        if self._info_fetch is not None:
            fetch, self._info_fetch = self._info_fetch, None
            fetch()

    @property
    def niter(self):
        self._info()
        return self._niter

    @property
    def residual(self):
        self._info()
        return self._residual

    @property
    def lam(self):
        self._info()
        return self._lam

    def __len__(self):
        return len(self.done)

    def terminal_observations(self):
        if self._terminal is None:
            self._terminal = self._terminal_fetch()
        return self._terminal

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        d = {"residual": self.residual[i], "niter": int(self.niter[i]), "lam": complex(self.lam[i])}
        if self.done[i]:
            if self._truncated_key[i]:
                d["TimeLimit.truncated"] = False
            d["terminal_observation"] = self.terminal_observations()[i]
        return d


class _HostSet:
    """One page-locked host result block (layout ``sdcgym_block_layout``, include/sdcgym.h) and the numpy arrays a
    ``step`` hands out as views of it.  The observation is the strided view (N, 2, M) complex128 over the u rows and
    the residual rows of the block.

    ``free()`` tells whether the caller still holds any of the arrays handed out (or a view derived from them): every
    such view keeps a reference to ``root`` or to the array itself, so the reference counts are back at their
    construction-time values exactly when nothing outside this object can observe the memory any more."""

    def __init__(self, torch, layout, N, M, const_u):
        self.blk = torch.zeros(int(layout.total), dtype=torch.uint8, pin_memory=True)
        self.ptr = self.blk.data_ptr()
        root = self.blk.numpy()
        self.root = root

        def seg(off, nbytes, dtype):
            return root[int(off): int(off) + nbytes].view(dtype)

        u_rows = seg(layout.obs_u, N * M * 16, np.complex128)
        if const_u:
            u_rows[:] = 1.0  # sdc-v0 + auto-reset: the returned u is the reset state, filled once, never transferred
        self.obs = np.lib.stride_tricks.as_strided(u_rows, shape=(N, 2, M),
                                                   strides=(16 * M, int(layout.obs_r - layout.obs_u), 16))
        self.reward = seg(layout.reward, N * 8, np.float64)
        self.residual = seg(layout.residual, N * 8, np.float64)
        self.lam = seg(layout.lam, N * 16, np.complex128)
        self.niter = seg(layout.niter, N * 4, np.int32)
        self.flags = seg(layout.flags, N, np.uint8)
        self._done_u8 = np.zeros(N, np.uint8)  # `dones`, derived from the flags on the host (not part of the block)
        self.dones = self._done_u8.view(np.bool_)
        del u_rows, root
        self._handed = (self.obs, self.reward, self.residual, self.lam, self.niter, self.flags, self._done_u8, self.dones)
        self._baseline = self._counts()

    def _counts(self):
        return [sys.getrefcount(self.root)] + [sys.getrefcount(h) for h in self._handed]

    def free(self):
        return self._counts() == self._baseline


class _TruncatedKey:
    """``'TimeLimit.truncated' in info`` per env: gym's TimeLimit adds the key once the episode has run
    ``max_episode_steps`` steps (always for sdc-v0, at niter >= 50 for sdc-v1).  Evaluated lazily."""

    def __init__(self, niter, max_steps):
        self._niter, self._max = niter, max_steps  # `niter`: an array, or a callable returning it (lazy info arrays)

    def __getitem__(self, i):
        n = self._niter() if callable(self._niter) else self._niter
        return bool(n[i] >= self._max)


class _EnvProxy:
    """``vec_env.envs[i]``: attribute view of one env of the batch (reads go through a cached host snapshot)."""

    def __init__(self, vec, i):
        self._vec, self._i = vec, i

    prec = property(lambda self: self._vec.prec)
    restol = property(lambda self: self._vec.restol)
    M = property(lambda self: self._vec.M)
    dt = property(lambda self: self._vec.dt)
    Q = property(lambda self: self._vec.Q)
    max_iters = MAX_ITERS
    action_space = property(lambda self: self._vec.action_space)
    observation_space = property(lambda self: self._vec.observation_space)

    @property
    def lam(self):
        return complex(self._vec._snapshot()["lam"][self._i])

    @property
    def state(self):
        obs = self._vec._snapshot()["obs"][self._i]
        return (obs[0], obs[1])

    @property
    def niter(self):
        return int(self._vec._snapshot()["niter"][self._i])

    @property
    def num_episodes(self):
        return int(self._vec._snapshot()["episodes"][self._i])

    @property
    def initial_residual(self):
        # a function of lambda only: r0 = u0 - C @ 1 (sdc_env.py:311-313); recomputed on the device
        return self._vec._initial_residuals()[self._i]

    def set_num_episodes(self, n):
        self._vec.set_num_episodes(n, indices=[self._i])


class _EnvProxies(Sequence):
    def __init__(self, vec):
        self._vec = vec

    def __len__(self):
        return self._vec.num_envs

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return _EnvProxy(self._vec, i)


class SDCVecEnv:
    def __init__(
        self,
        envname: str = "sdc-v0",
        num_envs: int = 1,
        M: Optional[int] = None,
        dt: Optional[float] = None,
        restol: Optional[float] = None,
        prec: Optional[str] = None,
        seed: Optional[int] = None,
        lambda_real_interval=(-100, 0),
        lambda_imag_interval=(0, 0),
        lambda_real_interpolation_interval=None,
        norm_factor=1,
        residual_weight=0.5,
        step_penalty=0.1,
        reward_iteration_only=None,
        reward_strategy="iteration_only",
        collect_states=False,
        use_doubles=True,
        do_scale=True,
        free_action_space=False,
        # ---- extensions over the reference constructor ----
        prec_type: str = "diag",
        Q: Optional[np.ndarray] = None,
        device=None,
        env_offset: int = 0,
        blas_variant: Optional[int] = None,
        autoreset: bool = True,
        output: str = "numpy",
        reuse_buffers: bool = False,
        pipeline_chunks: int = 0,
        host_pipeline: str = "native",
        max_host_sets: int = 4,
        keep_terminal: bool = True,
        sweep_mode: str = "exact",
        lazy_info: bool = True,
        phased: Optional[bool] = None,
    ):
        torch = _torch()
        if envname not in _lib.ENV_KINDS:
            raise ValueError(f"unknown env id {envname!r} (supported: sdc-v0, sdc-v1)")
        if M is None or dt is None or restol is None:
            raise TypeError("M, dt and restol are required (as in the reference constructor)")
        if not torch.cuda.is_available():
            raise _lib.SdcGymError("SDCVecEnv needs a CUDA device: the SDC kernels have no CPU fallback")
        # use_doubles=False (SAC, utils/utils.py:279-280): float32 / complex64 action space; the reference then stores
        # Q_delta in that dtype (sdc_env.py:138-140) and sweeps in complex128 - the kernels round the scaled action to
        # float32 (SDCGYM_ACTION_F32) and compute in fp64 as before
        self.use_doubles = bool(use_doubles)
        self._L = _lib.load()
        self.envname = envname
        self.num_envs = int(num_envs)
        self.M, self.dt, self.restol = int(M), float(dt), float(restol)
        self.prec = prec
        self.prec_type = "fixed" if prec is not None else prec_type
        if not self._L.sdcgym_supported(self.M, _lib.PREC_TYPES[self.prec_type]):
            raise NotImplementedError(f"M={self.M}, prec_type={self.prec_type} has no compiled kernel")
        self.Q = collocation_matrix(self.M) if Q is None else np.ascontiguousarray(Q, dtype=np.float64)
        self.Qd_fixed = fixed_preconditioner(prec, self.M, self.Q) if prec is not None else np.zeros((self.M, self.M))
        self.lambda_real_interval = list(lambda_real_interval)
        self.lambda_imag_interval = list(lambda_imag_interval)
        self.lambda_real_interpolation_interval = lambda_real_interpolation_interval
        self.norm_factor, self.residual_weight, self.step_penalty = norm_factor, residual_weight, step_penalty
        if reward_iteration_only is None:
            self.reward_strategy = reward_strategy.lower()
        elif reward_iteration_only:
            self.reward_strategy = "iteration_only"
        else:
            self.reward_strategy = "residual_change"
        if self.reward_strategy not in _lib.REWARD_STRATEGIES:
            raise NotImplementedError(f"unknown reward strategy {self.reward_strategy}")
        self.collect_states = bool(collect_states)
        self.do_scale = bool(do_scale)
        self.free_action_space = bool(free_action_space)
        self.autoreset = bool(autoreset)
        self.output = output
        self.reuse_buffers = bool(reuse_buffers)
        self.max_iters = MAX_ITERS
        self.n_act = num_actions(self.M, prec_type) if prec is None else self.M
        self._kernel_n_act = 0 if prec is not None else self.n_act

        # spaces (sdc_env.py:89-110)
        obs_shape = (self.M * 2, self.max_iters) if collect_states else (2, self.M)
        self.observation_space = Box(-1e10, 1e10, obs_shape, np.complex128)
        if free_action_space:
            self.action_space = Box(-np.inf, np.inf, (self.n_act,), np.complex128 if use_doubles else np.complex64)
        else:
            self.action_space = Box(-1.0, 1.0, (self.n_act,), np.float64 if use_doubles else np.float32)

        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        N = self.num_envs
        self.ld = max(32, (N + 31) // 32 * 32)
        f64, dev = torch.float64, self.device
        with torch.cuda.device(self.device):
            self.lam = torch.zeros((2, self.ld), dtype=f64, device=dev)
            self.S = torch.zeros((4 * self.M, self.ld), dtype=f64, device=dev)
            self.resnorm = torch.zeros(self.ld, dtype=f64, device=dev)
            self.niter = torch.zeros(self.ld, dtype=torch.int32, device=dev)
            self.episodes = torch.zeros(self.ld, dtype=torch.int32, device=dev)
            self.rng_ctr = torch.zeros(self.ld, dtype=torch.int32, device=dev)
            # ||initial residual|| of the running episode, kept for the residual_change reward (sdcgym_state.norm_init)
            self.norm_init = torch.zeros(self.ld, dtype=f64, device=dev)
            # results of a step: ONE device block (include/sdcgym.h: sdcgym_block_layout) whose host twin is what
            # `step` hands out, so a step's results leave the GPU in a single transfer; the per-array tensors below
            # are views of it
            self._layout = _lib.BlockLayout()
            _lib.check(self._L.sdcgym_block_layout_init(self.M, N, ctypes.byref(self._layout)), "sdcgym_block_layout_init")
            lay = self._layout
            self.dev_block = torch.zeros(max(1, int(lay.total)), dtype=torch.uint8, device=dev)

            def seg(off, nbytes, dtype):
                return self.dev_block[int(off): int(off) + nbytes].view(dtype)

            self.reward = seg(lay.reward, N * 8, f64)
            self.flags = seg(lay.flags, N, torch.uint8)
            self.info_residual = seg(lay.residual, N * 8, f64)
            self.info_niter = seg(lay.niter, N * 4, torch.int32)
            self.info_lam = seg(lay.lam, N * 16, f64).view(N, 2)
            self.terminal = torch.zeros((4 * self.M, self.ld), dtype=f64, device=dev) if keep_terminal else None
            self.obs_aos = torch.zeros((N, 2, self.M, 2), dtype=f64, device=dev)
            a_w = max(1, self._kernel_n_act) * (2 if free_action_space else 1)
            self.action_dev = torch.zeros((N, a_w), dtype=f64, device=dev)
            self.old_states = (torch.zeros((N, 2 * self.M, self.max_iters, 2), dtype=f64, device=dev)
                               if collect_states else None)
            # sweep_mode='certified' (sdc-v0): substitution sweeps + per-env certificate, exact kernel for the envs whose
            # decisions fall inside the error bound (include/sdcgym.h SDCGYM_SWEEP_CERTIFIED)
            if sweep_mode not in _lib.SWEEP_MODES:
                raise ValueError(f"sweep_mode must be one of {sorted(_lib.SWEEP_MODES)}")
            self.sweep_mode = sweep_mode
            self._certified = sweep_mode == "certified" and envname == "sdc-v0" and not collect_states
            if self._certified:
                self.cert = torch.zeros((_lib.CERT_PLANES, self.ld), dtype=torch.float32, device=dev)
                self.fallback_list = torch.zeros(max(1, N), dtype=torch.int32, device=dev)
                self.fallback_count = torch.zeros(2, dtype=torch.int32, device=dev)
            # phased full solve (include/sdcgym.h: sdcgym_state.phase_*): sdc-v0 with a non-diagonal Q_delta - the envs
            # of a warp stop after very different sweep counts, so large batches are solved in passes over compacted
            # lists of the envs still iterating.  Bit-identical to the single launch; `phased=False` keeps that one.
            dense = self.prec_type != "diag"
            if prec is not None:
                dense = bool(np.any(self.Qd_fixed != np.diag(np.diag(self.Qd_fixed))))
            #   phased=None (default): large batches allocate the work buffers and `step_tensor` TIMES both launch
            #   sequences on the caller's workload (CUDA events, no synchronisation) and keeps the faster one - the gain
            #   depends on how unevenly the envs of a warp finish (measured: 1.5x when most envs stop after a few sweeps,
            #   a few per cent slower when a third of them run all 50 sweeps; profiles/README.md);
            #   phased=True: always the phased sequence.
            want = (envname == "sdc-v0" and dense and self.M <= PHASED_MAX_M and not collect_states
                    and N >= PHASED_MIN_ENVS) if phased is None else bool(phased)
            if want and phased is None:
                # the work planes (2 M^2 doubles + two list entries per env) are only worth a quarter of the free memory
                need = (2 * self.M * self.M * 8 + 8) * self.ld
                want = need <= torch.cuda.mem_get_info(dev)[0] // 4
            self.phased = bool(want and envname == "sdc-v0" and dense and self.M <= PHASED_MAX_M and not collect_states)
            self._phase_auto = self.phased and phased is None
            self._phase_use = self.phased  # what the next device step launches
            self._phase_trial = None       # running A/B measurement: {"ev": {True: [...], False: [...]}, "k": launches so far}
            self._phase_next_trial = 0     # step count at which the next measurement starts
            self.phase_timings = None      # (phased ms, single-launch ms) of the last completed measurement
            if self.phased:
                self.phase_list = torch.zeros(2 * max(1, N), dtype=torch.int32, device=dev)
                self.phase_count = torch.zeros(_lib.PHASE_COUNTERS, dtype=torch.int32, device=dev)
                self.phase_pinv = torch.empty((2 * self.M * self.M, self.ld), dtype=f64, device=dev)
        self._state_exact = True  # the stored (u, r) are bit-equal to the reference's (reset / exact step / injected)
        self._host = None  # pinned staging, allocated on first numpy-mode step
        self._snap = None
        self._init_res = None
        self._pending = None
        self.pipeline_chunks = int(pipeline_chunks)
        if host_pipeline != "native":
            raise ValueError("host_pipeline: only 'native' (sdcgym_pipe_step_block inside libsdcgym.so) exists")
        self.host_pipeline = host_pipeline
        self.max_host_sets = max(1, int(max_host_sets))
        self.keep_terminal = bool(keep_terminal)
        self.lazy_info = bool(lazy_info)
        self._pipe = None
        self._bio = None
        self._step_count = 0
        self._term_aos = None
        self.host_set_copies = 0  # steps that had to copy their outputs (all host sets still referenced by the caller)

        self._desc = _lib.EnvDesc()
        d = self._desc
        d.M, d.env_kind = self.M, _lib.ENV_KINDS[envname]
        d.prec_type = _lib.PREC_TYPES[self.prec_type]
        d.action_is_complex = int(self.free_action_space)
        d.do_scale = (_lib.ACTION_SCALE if self.do_scale else 0) | (0 if self.use_doubles else _lib.ACTION_F32)
        d.max_iters = self.max_iters
        # 'spectral_radius' (sdc_env.py:421-425) is composed from two kernels: the step runs with the cheapest
        # in-kernel reward and `_apply_spectral_radius_reward` overwrites it with rho from the eigenvalue kernel
        self._rho_reward = self.reward_strategy == "spectral_radius"
        if self._rho_reward and not self.use_doubles and prec is None:
            raise NotImplementedError("reward_strategy='spectral_radius' with use_doubles=False: the eigenvalue kernel "
                                      "forms P in complex128, the reference's float32 Q_delta path does not")
        d.reward_strategy = _lib.REWARD_STRATEGIES["iteration_only" if self._rho_reward else self.reward_strategy]
        d.blas_variant = detect_blas_variant() if blas_variant is None else int(blas_variant)
        d.autoreset = int(self.autoreset and not self.collect_states)
        d.curriculum = int(lambda_real_interpolation_interval is not None)
        d.sweep_mode = _lib.SWEEP_MODES["certified" if self._certified else "exact"]
        d.dt, d.restol = self.dt, self.restol
        d.step_penalty, d.residual_weight, d.norm_factor = float(step_penalty), float(residual_weight), float(norm_factor)
        d.lam_re_lo, d.lam_re_hi = float(lambda_real_interval[0]), float(lambda_real_interval[1])
        d.lam_im_lo, d.lam_im_hi = float(lambda_imag_interval[0]), float(lambda_imag_interval[1])
        if lambda_real_interpolation_interval is not None:
            d.interp_x0, d.interp_x1 = (float(v) for v in lambda_real_interpolation_interval)
        # seed=None: fresh OS entropy like the reference's gym seeding (every env built without a seed differs);
        # the key actually used is kept in `seed_used` so a run can be reproduced
        self.seed_used = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        d.seed = self.seed_used
        d.env_offset = int(env_offset)
        for k, v in enumerate(self.Q.reshape(-1)):
            d.Q[k] = float(v)
        for k, v in enumerate(np.asarray(self.Qd_fixed, dtype=np.float64).reshape(-1)):
            d.Qd_fixed[k] = float(v)
        self.blas_variant = d.blas_variant
        self.envs = _EnvProxies(self)
        # sdc-v0 with the fused auto-reset: every step ends the episode, the returned observation is the reset state
        # of the next lambda, whose u row is identically 1 (sdc_env.py:306-314) - never exported, never transferred
        self._const_u = bool(envname == "sdc-v0" and d.autoreset)
        self._guard = _lib.DeviceGuard(self.device)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _state(self, start=0, count=None, use_phased=True):
        count = self.num_envs - start if count is None else count
        st = _lib.State()
        st.N, st.ld = count, self.ld
        st.lam = self.lam.data_ptr() + 8 * start
        st.S = self.S.data_ptr() + 8 * start
        st.resnorm = self.resnorm.data_ptr() + 8 * start
        st.niter = self.niter.data_ptr() + 4 * start
        st.episodes = self.episodes.data_ptr() + 4 * start
        st.rng_ctr = self.rng_ctr.data_ptr() + 4 * start
        st.norm_init = self.norm_init.data_ptr() + 8 * start
        if self._certified:
            st.cert = self.cert.data_ptr() + 4 * start
            st.fallback_list = self.fallback_list.data_ptr() + 4 * start
            st.fallback_count = self.fallback_count.data_ptr()
        if self.phased and use_phased:
            st.phase_list = self.phase_list.data_ptr() + 8 * start
            st.phase_count = self.phase_count.data_ptr()
            st.phase_pinv = self.phase_pinv.data_ptr() + 8 * start
        return st

    def _desc_for(self, start):
        if start == 0:
            return self._desc
        d = _lib.EnvDesc()
        ctypes.memmove(ctypes.byref(d), ctypes.byref(self._desc), ctypes.sizeof(d))
        d.env_offset = self._desc.env_offset + start
        return d

    def _invalidate(self):
        self._snap = None
        self._init_res = None

    def _certified_step_begins(self):
        """The certificate of sweep_mode='certified' compares against a reference that starts from the SAME bits: a
        step may only start from an exact state (reset, auto-reset, set_state), never from the rounding-level
        approximation a previous certified step left behind (possible only with autoreset=False)."""
        if not self._certified:
            return
        if not self._state_exact:
            raise _lib.SdcGymError("sweep_mode='certified': this step would start from the result of a previous "
                                   "certified step (autoreset=False); call reset() / set_state() first or use "
                                   "sweep_mode='exact'")
        if not self._desc.autoreset:
            self._state_exact = False

    def fallback_stats(self):
        """sweep_mode='certified': (envs re-run by the exact kernel in the last step [last chunk of a pipelined host
        step], cumulative number over all steps).  Synchronises."""
        if not self._certified:
            return 0, 0
        c = self.fallback_count.cpu().numpy()
        return int(c[0]), int(c[1])

    # ------------------------------------------------------------------ reset / seed
    @_lib.on_device
    def seed(self, seed=None):
        """DummyVecEnv.seed: env i gets ``seed + i`` in the reference; here one Philox key for the batch."""
        self.seed_used = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self._desc.seed = self.seed_used
        self.rng_ctr.zero_()
        # DummyVecEnv returns the per-env seeds; a lazy range instead of a list of num_envs integers
        return [None] * self.num_envs if seed is None else range(int(seed), int(seed) + self.num_envs)

    @_lib.on_device
    def set_num_episodes(self, num_episodes, indices=None):
        if indices is None:
            self.episodes.fill_(int(num_episodes))
        else:
            idx = _torch().as_tensor(list(indices), device=self.device, dtype=_torch().long)
            self.episodes[idx] = int(num_episodes)
        self._invalidate()

    @_lib.on_device
    def reset(self, lam=None, mask=None):
        """Reset all envs (or those with ``mask``); ``lam`` (N,) complex injects the lambdas instead of drawing."""
        torch = _torch()
        lam_ptr = None
        if lam is not None:
            lam_t = torch.as_tensor(np.ascontiguousarray(np.asarray(lam, dtype=np.complex128).reshape(-1)))
            if lam_t.numel() != self.num_envs:
                raise ValueError("lam must have one entry per env")
            lam_planes = torch.zeros((2, self.ld), dtype=torch.float64, device=self.device)
            ri = torch.view_as_real(lam_t).to(self.device)
            lam_planes[0, : self.num_envs] = ri[:, 0]
            lam_planes[1, : self.num_envs] = ri[:, 1]
            lam_ptr = lam_planes.data_ptr()
        mask_ptr = None
        if mask is not None:
            mask_t = torch.as_tensor(np.asarray(mask, dtype=np.uint8)).to(self.device)
            mask_ptr = mask_t.data_ptr()
        st = self._state()
        os_ptr = self.old_states.data_ptr() if self.old_states is not None else None
        _lib.check(self._L.sdcgym_reset(ctypes.byref(self._desc), ctypes.byref(st), lam_ptr, mask_ptr, os_ptr,
                                        self._stream()), "sdcgym_reset")
        if mask is None:
            self._state_exact = True
        self._invalidate()
        return self._observation()

    # ------------------------------------------------------------------ observations
    def _export_obs(self, src, start=0, count=None):
        count = self.num_envs - start if count is None else count
        _lib.check(self._L.sdcgym_export_obs(self.M, count, self.ld, src.data_ptr() + 8 * start,
                                             self.obs_aos.data_ptr() + 8 * start * 4 * self.M, self._stream()),
                   "sdcgym_export_obs")

    @_lib.on_device
    def observation_tensor(self):
        """Current observation as a CUDA complex128 tensor (N, 2, M) (or (N, 2M, 50) with collect_states)."""
        torch = _torch()
        if self.collect_states:
            return torch.view_as_complex(self.old_states)
        self._export_obs(self.S)
        return torch.view_as_complex(self.obs_aos)

    def _observation(self):
        t = self.observation_tensor()
        if self.output == "torch":
            return t
        return t.cpu().numpy()

    @_lib.on_device
    def _snapshot(self):
        if self._snap is None:
            torch = _torch()
            N = self.num_envs
            self._export_obs(self.S)
            obs = torch.view_as_complex(self.obs_aos).cpu().numpy()
            lam = self.lam[:, :N].cpu().numpy()
            self._snap = dict(obs=obs, lam=lam[0] + 1j * lam[1], niter=self.niter[:N].cpu().numpy(),
                              episodes=self.episodes[:N].cpu().numpy())
        return self._snap

    @_lib.on_device
    def _rebuild_norm_init(self):
        """norm_init of the running episodes from their lambdas (reset kernel on scratch planes)."""
        torch = _torch()
        scratch = dict(lam=torch.empty_like(self.lam), S=torch.empty_like(self.S), resnorm=torch.empty_like(self.resnorm),
                       niter=torch.empty_like(self.niter), episodes=torch.zeros_like(self.episodes),
                       rng_ctr=torch.zeros_like(self.rng_ctr))
        st = _lib.State()
        st.N, st.ld = self.num_envs, self.ld
        for k, t in scratch.items():
            setattr(st, k, t.data_ptr())
        st.norm_init = self.norm_init.data_ptr()
        _lib.check(self._L.sdcgym_reset(ctypes.byref(self._desc), ctypes.byref(st), self.lam.data_ptr(), None, None,
                                        self._stream()), "sdcgym_reset")
        _torch().cuda.current_stream(self.device).synchronize()

    def _initial_residuals(self):
        if self._init_res is None:
            # r0 is a function of lambda only: run the reset kernel on scratch planes with lambda injected
            torch = _torch()
            N, M = self.num_envs, self.M
            scratch = dict(lam=torch.empty_like(self.lam), S=torch.empty_like(self.S),
                           resnorm=torch.empty_like(self.resnorm), niter=torch.empty_like(self.niter),
                           episodes=torch.zeros_like(self.episodes), rng_ctr=torch.zeros_like(self.rng_ctr))
            st = _lib.State()
            st.N, st.ld = N, self.ld
            for k, t in scratch.items():
                setattr(st, k, t.data_ptr())
            _lib.check(self._L.sdcgym_reset(ctypes.byref(self._desc), ctypes.byref(st), self.lam.data_ptr(), None, None,
                                            self._stream()), "sdcgym_reset")
            r = scratch["S"][2 * M:, :N].cpu().numpy()
            self._init_res = (r[0::2] + 1j * r[1::2]).T.copy()
        return self._init_res

    # ------------------------------------------------------------------ step
    def _launch_step(self, action_ptr, env_stride, comp_stride, start=0, count=None, want_terminal=True):
        count = self.num_envs - start if count is None else count
        io = _lib.StepIO()
        io.action = action_ptr
        io.action_env_stride, io.action_comp_stride = env_stride, comp_stride
        io.reward = self.reward.data_ptr() + 8 * start
        io.flags = self.flags.data_ptr() + start
        io.info_residual = self.info_residual.data_ptr() + 8 * start
        io.info_niter = self.info_niter.data_ptr() + 4 * start
        io.info_lam = self.info_lam.data_ptr() + 16 * start
        io.terminal_obs = (self.terminal.data_ptr() + 8 * start) if (want_terminal and self.terminal is not None) else None
        io.old_states = (self.old_states.data_ptr() + 8 * start * 2 * self.M * self.max_iters * 2
                         if self.old_states is not None else None)
        whole = self.phased and start == 0 and count == self.num_envs
        use_phased, ev = (self._phase_pick() if whole and self._phase_auto else (self.phased, None))
        st = self._state(start, count, use_phased)
        d = self._desc_for(start)
        self._certified_step_begins()
        if ev is not None:
            ev[0].record(_torch().cuda.current_stream(self.device))
        _lib.check(self._L.sdcgym_step(ctypes.byref(d), ctypes.byref(st), ctypes.byref(io), self._stream()),
                   "sdcgym_step")
        if ev is not None:
            ev[1].record(_torch().cuda.current_stream(self.device))

    def _phase_pick(self):
        """phased=None: which launch sequence this device step takes, and the event pair that times it (or None).

        A measurement is PHASED_TRIAL launches of each sequence, alternating, the first of each untimed; it is read
        back without blocking once all its events have completed, and repeated every PHASED_RETUNE steps (the
        workload of a learning policy drifts).  Results are bit-identical either way: only the speed is chosen."""
        torch = _torch()
        if torch.cuda.is_current_stream_capturing():  # inside a CUDA graph capture: no events, no queries
            return self._phase_use, None
        t = self._phase_trial
        if t is None and self._step_count >= self._phase_next_trial:
            t = self._phase_trial = {"ev": {True: [], False: []}, "k": 0}
        if t is None:
            return self._phase_use, None
        if t["k"] < 2 * PHASED_TRIAL:
            which = (t["k"] % 2 == 0)
            timed = t["k"] >= 2
            t["k"] += 1
            if not timed:
                return which, None
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            t["ev"][which].append(ev)
            return which, ev
        if all(e[1].query() for evs in t["ev"].values() for e in evs):
            ms = {w: min(e[0].elapsed_time(e[1]) for e in evs) for w, evs in t["ev"].items()}
            self.phase_timings = (ms[True], ms[False])
            self._phase_use = ms[True] < ms[False]
            self._phase_trial = None
            self._phase_next_trial = self._step_count + PHASED_RETUNE
        return self._phase_use, None

    @_lib.on_device
    def step_tensor(self, actions=None, want_terminal=True):
        """Device-resident step: ``actions`` is a CUDA float64 (N, A) / complex128 (N, A) tensor (ignored for
        fixed ``prec``).  Nothing is copied to the host and nothing synchronises.  Returns a dict of CUDA
        tensors (views on the env's own buffers, overwritten by the next step)."""
        torch = _torch()
        N = self.num_envs
        ptr, es, cs = None, 0, 0
        if self._kernel_n_act > 0:
            if actions is None:
                raise ValueError("actions required")
            if actions.is_complex():
                if not self.free_action_space:
                    raise TypeError("complex actions need free_action_space=True")
                a = torch.view_as_real(actions.to(torch.complex128))
                es, cs = a.stride(0), a.stride(1)
            else:
                if self.free_action_space:
                    a = torch.view_as_real(actions.to(torch.complex128))
                    es, cs = a.stride(0), a.stride(1)
                else:
                    a = actions.to(torch.float64)
                    es, cs = a.stride(0), a.stride(1)
            if a.shape[0] != N or a.shape[1] != self._kernel_n_act:
                raise ValueError(f"actions must have shape ({N}, {self._kernel_n_act})")
            self._keep = a
            ptr = a.data_ptr()
        self._launch_step(ptr, es, cs, want_terminal=want_terminal)
        if self._rho_reward:
            self._apply_spectral_radius_reward(actions if self._kernel_n_act else None)
        if self.collect_states and self.autoreset:
            self._reset_done_envs()
        self._step_count += 1
        self._invalidate()
        out = dict(reward=self.reward[:N], flags=self.flags[:N], niter=self.info_niter[:N],
                   residual=self.info_residual[:N], lam=torch.view_as_complex(self.info_lam)[:N])
        if self.terminal is not None:
            out["terminal"] = self.terminal[:, :N]
        return out

    def _apply_spectral_radius_reward(self, actions_t):
        """reward = rho(lam*dt*Pinv (Q - Qd)) for every env that did not err (reference reward_func :459-460)."""
        torch = _torch()
        from .loss import SpectralRadiusLoss

        if getattr(self, "_rho_loss", None) is None:
            self._rho_loss = SpectralRadiusLoss(self.M, self.dt, self.prec_type if self.prec is None else "diag",
                                                prec=self.prec, Q=self.Q, device=self.device)
        N = self.num_envs
        lam = torch.view_as_complex(self.info_lam)[:N]
        out = None
        if self.prec is None:
            out = actions_t
            if not self.free_action_space and self.do_scale:
                out = ((actions_t + 1.0) * 0.5).clamp(0.0, 1.0)  # _scale_action, sdc_env.py:125-132
        rho = self._rho_loss.spectral_radii(lam, out)
        err = (self.flags[:N] & _lib.FLAG_ERR).ne(0)
        self.reward[:N] = torch.where(err, self.reward[:N], rho)

    def _reset_done_envs(self):
        # collect_states: the kernel leaves finished envs alone; snapshot their buffers, then masked reset
        torch = _torch()
        N = self.num_envs
        done = (self.flags[:N] & _lib.FLAG_DONE).to(torch.uint8)
        self._terminal_old_states = torch.view_as_complex(self.old_states).clone()
        st = self._state()
        _lib.check(self._L.sdcgym_reset(ctypes.byref(self._desc), ctypes.byref(st), None, done.data_ptr(),
                                        self.old_states.data_ptr(), self._stream()), "sdcgym_reset")

    def _ensure_host(self):
        if self._host is None:
            torch = _torch()
            act = torch.zeros((self.num_envs, self.action_dev.shape[1]), dtype=torch.float64, pin_memory=True)
            # numpy views / pointers are built once: tensor.numpy() per step would cost more than a small batch's kernels
            self._host = dict(actions=[act], action_np=[act.numpy()], action_ptr=[act.data_ptr()], sets=[], spill=None)
            # page-locked allocation costs ~0.5 ms per MB: take the blocks the usual loop ping-pongs between now, not
            # inside somebody's second step
            for _ in range(1 if self.reuse_buffers else min(2, self.max_host_sets)):
                self._host["sets"].append(self._new_set())
        return self._host

    def _new_set(self):
        return _HostSet(_torch(), self._layout, self.num_envs, self.M, self._const_u)

    def _acquire_set(self, host):
        """(result set, owned): a page-locked result block that nothing outside the env references any more.

        ``DummyVecEnv.step`` returns arrays the caller owns; copying 123 MB per step (2^20 envs) to honour that costs
        five times the transfer itself.  Instead the step writes into a block whose arrays the caller has let go of
        (reference counts, ``_HostSet.free``) - a loop that rebinds ``obs, rew, done, info = env.step(a)`` ping-pongs
        between two blocks; a caller that keeps every result alive gets up to ``max_host_sets`` blocks and real copies
        after that (``host_set_copies`` counts those steps).  ``reuse_buffers=True`` always uses block 0."""
        sets = host["sets"]
        if self.reuse_buffers:
            if not sets:
                sets.append(self._new_set())
            return sets[0], True
        for hs in sets:
            if hs.free():
                return hs, True
        if len(sets) < self.max_host_sets:
            sets.append(self._new_set())
            return sets[-1], True
        if host["spill"] is None:
            host["spill"] = self._new_set()
        return host["spill"], False

    def pinned_action_buffer(self, index=0):
        """Page-locked numpy view (N, A) [(N, A) complex128 with free_action_space] that ``step`` uploads from
        without an intermediate copy: write the actions here and pass this very array to ``step``.  Two buffers
        (index 0 / 1) exist so a caller can fill one while the other is in flight."""
        host = self._ensure_host()
        while len(host["actions"]) <= index:
            host["actions"].append(_torch().zeros_like(host["actions"][0]).pin_memory())
            host["action_np"].append(host["actions"][-1].numpy())
            host["action_ptr"].append(host["actions"][-1].data_ptr())
        a = host["actions"][index].numpy()
        if self.free_action_space:
            return a.view(np.complex128)
        return a

    def _stage_actions(self, host, actions):
        """Return the pinned tensor holding ``actions`` (uploading from the caller's array when it already is one
        of ours, else copying it into staging buffer 0)."""
        if self.free_action_space:
            a = np.ascontiguousarray(actions, dtype=np.complex128).reshape(self.num_envs, self._kernel_n_act)
            a = a.view(np.float64)
        else:
            a = np.asarray(actions)
            if np.iscomplexobj(a):
                raise TypeError("complex actions need free_action_space=True")
            a = np.asarray(a, dtype=np.float64).reshape(self.num_envs, self._kernel_n_act)
        ptr = a.__array_interface__["data"][0]
        if a.flags.c_contiguous:
            for t, tp in zip(host["actions"], host["action_ptr"]):
                if tp == ptr:
                    return t
        np.copyto(host["action_np"][0], a)
        return host["actions"][0]

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        actions, self._pending = self._pending, None
        return self.step(actions)

    @_lib.on_device
    def step(self, actions):
        """(obs, rewards, dones, infos) like ``DummyVecEnv.step`` over the reference envs."""
        torch = _torch()
        if self.output == "torch" or (actions is not None and isinstance(actions, torch.Tensor) and actions.is_cuda):
            if self._kernel_n_act and not isinstance(actions, torch.Tensor):
                arr = np.asarray(actions, dtype=np.complex128 if self.free_action_space else np.float64)
                actions = torch.as_tensor(np.ascontiguousarray(arr.reshape(self.num_envs, self._kernel_n_act)))
            if self._kernel_n_act and not actions.is_cuda:
                actions = actions.to(self.device)
            out = self.step_tensor(actions if self._kernel_n_act else None)
            obs = self.observation_tensor()
            dones = (out["flags"] & _lib.FLAG_DONE).bool()
            return obs, out["reward"], dones, out
        if self._rho_reward:
            return self._step_simple(actions)
        if self.collect_states:
            host = self._ensure_host()
            src = self._stage_actions(host, actions) if self._kernel_n_act > 0 else None
            return self._step_collect_states(host, src)
        return self._step_host(actions)

    def _block_io(self):
        """The native host pipeline (csrc/hostpipe.cu) and its cached argument structs."""
        if self._pipe is None:
            handle = ctypes.c_void_p()
            _lib.check(self._L.sdcgym_pipe_create(64, ctypes.byref(handle)), "sdcgym_pipe_create")
            self._pipe = handle
        if self._bio is None:
            # the device buffers never move: build the argument structs once
            bio = _lib.BlockIO()
            bio.dev_block = self.dev_block.data_ptr()
            bio.action_dev = self.action_dev.data_ptr() if self._kernel_n_act else None
            bio.terminal_obs = self.terminal.data_ptr() if self.terminal is not None else None
            bio.skip_u = int(self._const_u) | (2 if self.lazy_info else 0)
            self._bio = (bio, self._state())
        self._bio[0].chunks = self.pipeline_chunks
        return self._bio

    def _step_host(self, actions, vn=None):
        """numpy in / numpy out as ONE C call (``sdcgym_pipe_step_block``): pinned actions -> H2D | step + export
        kernels | D2H straight into the result block whose views are returned."""
        host = self._ensure_host()
        src = self._stage_actions(host, actions) if self._kernel_n_act > 0 else None
        hs, owned = self._acquire_set(host)
        bio, st = self._block_io()
        if self.phased:  # phased=None: follow what the device steps measured (the struct is cached)
            on = self._phase_use
            st.phase_list = self.phase_list.data_ptr() if on else None
            st.phase_count = self.phase_count.data_ptr() if on else None
            st.phase_pinv = self.phase_pinv.data_ptr() if on else None
        bio.host_block = hs.ptr
        bio.action_host = src.data_ptr() if src is not None else None
        self._certified_step_begins()
        _lib.check(self._L.sdcgym_pipe_step_block(self._pipe, ctypes.byref(self._desc), ctypes.byref(st),
                                                  ctypes.byref(self._layout), ctypes.byref(bio),
                                                  None if vn is None else ctypes.byref(vn), self._stream()),
                   "sdcgym_pipe_step_block")
        self._step_count += 1
        self._invalidate()
        lazy = (self.lazy_info and vn is None
                and (self.pipeline_chunks > 1 or (self.pipeline_chunks <= 0 and self.num_envs >= 32768)))
        return self._host_outputs(hs, owned, lazy)

    def _step_simple(self, actions):
        """Unpipelined host step (rarely used configurations): upload, device step, download."""
        torch = _torch()
        N = self.num_envs
        a = None
        if self._kernel_n_act > 0:
            arr = np.asarray(actions, dtype=np.complex128 if self.free_action_space else np.float64)
            a = torch.as_tensor(np.ascontiguousarray(arr.reshape(N, self._kernel_n_act))).to(self.device)
        out = self.step_tensor(a)
        obs = self.observation_tensor().cpu().numpy()
        flags = out["flags"].cpu().numpy()
        dones = np.bitwise_and(flags, _lib.FLAG_DONE).view(np.bool_)
        niter = out["niter"].cpu().numpy()
        infos = LazyInfos(niter, out["residual"].cpu().numpy(), out["lam"].cpu().numpy(), dones,
                          _TruncatedKey(niter, MAX_EPISODE_STEPS[self.envname]), self._terminal_fetcher())
        infos.flags = flags
        return obs, out["reward"].cpu().numpy(), dones, infos

    def _host_outputs(self, hs, owned, lazy_info=False):
        if owned:
            cp = lambda x: x  # noqa: E731 - the block is the caller's until they drop it (see _acquire_set)
        else:
            self.host_set_copies += 1
            cp = np.ascontiguousarray if hs.obs.flags.c_contiguous else (lambda x: np.array(x, order="C"))
        if lazy_info and not owned:  # the copies below must see the info arrays
            self._fetch_info_into(hs)
            lazy_info = False
        obs, rewards, flags = cp(hs.obs), cp(hs.reward), cp(hs.flags)
        if owned:  # no per-step allocation: the mask lands in the block's own `dones` array
            np.bitwise_and(flags, _lib.FLAG_DONE, out=hs._done_u8)
            dones = hs.dones
        else:
            dones = np.bitwise_and(flags, _lib.FLAG_DONE).view(np.bool_)
        niter = cp(hs.niter)
        fetch = None
        if lazy_info:
            stamp = self._step_count

            def fetch():
                if self._step_count != stamp:
                    raise RuntimeError("info['niter' / 'residual' / 'lam'] of this step are gone: the env has stepped "
                                       "since (large batches keep them on the device until they are read; read them "
                                       "before the next step or construct the env with lazy_info=False)")
                self._fetch_info_into(hs)

        infos = LazyInfos(niter, cp(hs.residual), cp(hs.lam), dones, None, self._terminal_fetcher(), fetch)
        if lazy_info:  # (weak reference: a cycle would keep the result block referenced until the next gc run)
            ref = weakref.ref(infos)
            infos._truncated_key = _TruncatedKey(lambda: ref().niter, MAX_EPISODE_STEPS[self.envname])
        else:
            infos._truncated_key = _TruncatedKey(niter, MAX_EPISODE_STEPS[self.envname])
        infos.flags = flags
        return obs, rewards, dones, infos

    def _fetch_info_into(self, hs):
        """niter / residual / lam of the last step: device block -> the same bytes of the host block ``hs``."""
        lay = self._layout
        lo, hi = int(lay.residual), int(lay.flags)  # residual, lam, niter are adjacent (sdcgym_block_layout_init)
        with self._guard:
            hs.blk[lo:hi].copy_(self.dev_block[lo:hi])

    def _terminal_fetcher(self):
        """``info['terminal_observation']`` is served lazily from the device's terminal planes, which the NEXT step
        overwrites: an ``infos`` object read after a later step raises instead of returning that step's data."""
        stamp = self._step_count

        def fetch():
            if self._step_count != stamp:
                raise RuntimeError("terminal observations of this step are gone: the env has stepped since "
                                   "(read info['terminal_observation'] / infos.terminal_observations() before the next step)")
            return self._fetch_terminal()

        return fetch

    def _fetch_terminal(self):
        """terminal observations (N, 2, M) complex128 of the last step (valid rows: finished envs)."""
        torch = _torch()
        if self.collect_states:
            return self._terminal_old_states.cpu().numpy()
        if self.terminal is None:
            raise RuntimeError("terminal observations are not kept (keep_terminal=False)")
        with self._guard:
            if self._term_aos is None:
                self._term_aos = torch.empty_like(self.obs_aos)
            _lib.check(self._L.sdcgym_export_obs(self.M, self.num_envs, self.ld, self.terminal.data_ptr(),
                                                 self._term_aos.data_ptr(), self._stream()), "sdcgym_export_obs")
            return torch.view_as_complex(self._term_aos).cpu().numpy()

    def _step_collect_states(self, host, src):
        torch = _torch()
        N = self.num_envs
        a_w = self.action_dev.shape[1]
        if self._kernel_n_act > 0:
            self.action_dev.copy_(src, non_blocking=True)
        self._launch_step(self.action_dev.data_ptr() if self._kernel_n_act else None, a_w,
                          2 if self.free_action_space else 1)
        if self.autoreset:
            self._reset_done_envs()
        self._step_count += 1
        self._invalidate()
        obs = torch.view_as_complex(self.old_states).cpu().numpy()
        flags = self.flags[:N].cpu().numpy()
        dones = (flags & _lib.FLAG_DONE).astype(bool)
        lam = torch.view_as_complex(self.info_lam)[:N].cpu().numpy()
        niter = self.info_niter[:N].cpu().numpy()
        truncated_key = _TruncatedKey(niter, MAX_EPISODE_STEPS[self.envname])
        infos = LazyInfos(niter, self.info_residual[:N].cpu().numpy(), lam, dones, truncated_key,
                          self._terminal_fetcher())
        infos.flags = flags
        return obs, self.reward[:N].cpu().numpy(), dones, infos

    # ------------------------------------------------------------------ state injection (tests, dp_playground)
    @_lib.on_device
    def set_state(self, u, r, niter=None):
        """Overwrite (u, r) of all envs from host arrays (N, M) complex128 (the reference's ``env.state = ...``)."""
        torch = _torch()
        N, M = self.num_envs, self.M
        obs = np.stack([np.asarray(u, dtype=np.complex128).reshape(N, M), np.asarray(r, dtype=np.complex128).reshape(N, M)], axis=1)
        t = torch.view_as_real(torch.as_tensor(np.ascontiguousarray(obs))).to(self.device).contiguous()
        _lib.check(self._L.sdcgym_import_obs(M, N, self.ld, t.data_ptr(), self.S.data_ptr(), self._stream()),
                   "sdcgym_import_obs")
        _lib.check(self._L.sdcgym_refresh_resnorm(M, N, self.ld, self.S.data_ptr(), self.resnorm.data_ptr(),
                                                  self._stream()), "sdcgym_refresh_resnorm")
        if niter is not None:
            self.niter[:N] = torch.as_tensor(np.asarray(niter, dtype=np.int32)).to(self.device)
        torch.cuda.current_stream(self.device).synchronize()  # `t` must outlive the kernels
        self._state_exact = True
        self._invalidate()

    # ------------------------------------------------------------------ VecEnv odds and ends
    def get_attr(self, name, indices=None):
        idx = range(self.num_envs) if indices is None else indices
        return [getattr(self.envs[i], name) for i in idx]

    def set_attr(self, name, value, indices=None):
        if name == "num_episodes":
            self.set_num_episodes(value, indices)
        else:
            raise AttributeError(f"cannot set {name!r} on a batched env")

    def env_method(self, name, *args, indices=None, **kwargs):
        if name == "set_num_episodes":
            self.set_num_episodes(*args, indices=indices, **kwargs)
            return [None] * (self.num_envs if indices is None else len(indices))
        raise AttributeError(name)

    @_lib.on_device
    def state_dict(self):
        """Checkpointable env state (device tensors cloned to host)."""
        N = self.num_envs
        return {k: getattr(self, k)[..., :N].cpu() for k in ("lam", "S", "resnorm", "niter", "episodes", "rng_ctr",
                                                             "norm_init")} | {
            "seed": int(self._desc.seed), "state_exact": bool(self._state_exact)}

    @_lib.on_device
    def load_state_dict(self, sd):
        N = self.num_envs
        for k in ("lam", "S", "resnorm", "niter", "episodes", "rng_ctr"):
            getattr(self, k)[..., :N].copy_(sd[k].to(self.device))
        if "norm_init" in sd:
            self.norm_init[:N].copy_(sd["norm_init"].to(self.device))
        else:  # a checkpoint written before the plane existed: re-derive it from the lambdas
            self._rebuild_norm_init()
        self._desc.seed = int(sd["seed"])
        # a checkpoint is a state the caller vouches for (like set_state): certified steps may start from it.  States
        # saved after a certified step WITHOUT auto-reset carry that step's rounding-level approximation.
        self._state_exact = bool(sd.get("state_exact", True))
        self._invalidate()

    @_lib.on_device
    def close(self):
        self._host = None
        self._bio = None
        if self._pipe is not None:
            self._L.sdcgym_pipe_destroy(self._pipe)
            self._pipe = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, *a, **k):  # pragma: no cover
        raise NotImplementedError
