"""TEST INFRASTRUCTURE ONLY: host build of the kernel templates (see shim.cpp) + a tiny planar-state driver."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from sdc_gym_b200 import _lib
from sdc_gym_b200.collocation import collocation_matrix
from sdc_gym_b200.precond import fixed_preconditioner, num_actions

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsdcgym_hostshim.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "sdc_gym_b200", "csrc")


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    srcs = [os.path.join(_HERE, "shim.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    srcs.append(os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "sdcgym.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build():
    if _stale():
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma",
                               "-o", _SO, os.path.join(_HERE, "shim.cpp")])
    return _SO


_shim = None


def shim():
    global _shim
    if _shim is None:
        L = ctypes.CDLL(build())
        vp = ctypes.c_void_p
        L.shim_reset.argtypes = [ctypes.POINTER(_lib.EnvDesc), ctypes.POINTER(_lib.State), vp, vp, vp]
        L.shim_step.argtypes = [ctypes.POINTER(_lib.EnvDesc), ctypes.POINTER(_lib.State), ctypes.POINTER(_lib.StepIO)]
        L.shim_step_hold5.argtypes = L.shim_step.argtypes
        L.shim_step_hold9.argtypes = L.shim_step.argtypes
        L.shim_spectral_radius.argtypes = [ctypes.POINTER(_lib.RhoDesc), ctypes.c_int64, vp, vp, vp]
        L.shim_step_certified.argtypes = L.shim_step.argtypes + [vp, ctypes.c_int]
        L.shim_step_phased.argtypes = L.shim_step.argtypes + [vp, ctypes.c_int, ctypes.c_int]
        _shim = L
    return _shim


def make_desc(kind, M, *, prec=None, prec_type="diag", dt=1.0, restol=1e-10, cplx=False, do_scale=True,
              strategy="iteration_only", step_penalty=0.1, residual_weight=0.5, norm_factor=1.0, variant=0,
              autoreset=False, Q=None, seed=0, env_offset=0, re_int=(-100, 0), im_int=(-10, 0), curriculum=None,
              use_doubles=True):
    d = _lib.EnvDesc()
    d.M, d.env_kind = M, _lib.ENV_KINDS[kind]
    d.prec_type = _lib.PREC_TYPES["fixed" if prec is not None else prec_type]
    d.action_is_complex, d.do_scale, d.max_iters = int(cplx), int(bool(do_scale)) | (0 if use_doubles else 2), 50
    d.reward_strategy = _lib.REWARD_STRATEGIES[strategy]
    d.blas_variant, d.autoreset = variant, int(autoreset)
    d.dt, d.restol = dt, restol
    d.step_penalty, d.residual_weight, d.norm_factor = step_penalty, residual_weight, float(norm_factor)
    d.lam_re_lo, d.lam_re_hi, d.lam_im_lo, d.lam_im_hi = re_int[0], re_int[1], im_int[0], im_int[1]
    if curriculum is not None:
        d.curriculum, d.interp_x0, d.interp_x1 = 1, curriculum[0], curriculum[1]
    d.seed, d.env_offset = seed, env_offset
    Q = collocation_matrix(M) if Q is None else np.asarray(Q, dtype=np.float64)
    for k, v in enumerate(Q.reshape(-1)):
        d.Q[k] = float(v)
    if prec is not None:
        for k, v in enumerate(fixed_preconditioner(prec, M, Q).reshape(-1)):
            d.Qd_fixed[k] = float(v)
    return d


class ShimBatch:
    """Planar state of n envs in host memory, stepped by the host-compiled kernel bodies."""

    def __init__(self, desc, n, collect=False, entry="shim_step"):
        self.d, self.n, self.M = desc, n, desc.M
        self.entry = entry
        self.ld = max(32, (n + 31) // 32 * 32)
        M, ld = self.M, self.ld
        self.lam = np.zeros((2, ld))
        self.S = np.zeros((4 * M, ld))
        self.resnorm = np.zeros(ld)
        self.niter = np.zeros(ld, np.int32)
        self.episodes = np.zeros(ld, np.int32)
        self.rng_ctr = np.zeros(ld, np.uint32)
        self.norm_init = np.zeros(ld)  # cached ||initial residual|| (sdcgym_state.norm_init); None: re-derived per step
        self.reward = np.zeros(ld)
        self.flags = np.zeros(ld, np.uint8)
        self.info_res = np.zeros(ld)
        self.info_niter = np.zeros(ld, np.int32)
        self.info_lam = np.zeros((ld, 2))
        self.term = np.zeros((4 * M, ld))
        self.old_states = np.zeros((n, 2 * M, 50), np.complex128) if collect else None
        # work buffers of the certified sweep mode
        self.cert = np.zeros((_lib.CERT_PLANES, ld), np.float32)
        self.fallback_list = np.zeros(max(1, n), np.int32)
        self.fallback_count = np.zeros(2, np.int32)
        self.trace = None
        # work buffers of the phased dense solve (poisoned: a pass may only read what an earlier pass wrote)
        self.phase_list = np.full(2 * max(1, n), -1, np.int32)
        self.phase_count = np.full(_lib.PHASE_COUNTERS, -1, np.int32)
        self.phase_pinv = np.full((2 * M * M, ld), np.nan)
        self.phase_stops = (6, 16)
        self.phase_split = True  # inverse-only pass first (the library's default)

    def _state(self):
        st = _lib.State()
        st.N, st.ld = self.n, self.ld
        for k in ("lam", "S", "resnorm", "niter", "episodes", "rng_ctr"):
            setattr(st, k, getattr(self, k).ctypes.data)
        st.norm_init = None if self.norm_init is None else self.norm_init.ctypes.data
        st.cert, st.fallback_list = self.cert.ctypes.data, self.fallback_list.ctypes.data
        st.fallback_count = self.fallback_count.ctypes.data
        st.phase_list, st.phase_count = self.phase_list.ctypes.data, self.phase_count.ctypes.data
        st.phase_pinv = self.phase_pinv.ctypes.data
        return st

    def reset(self, lam=None, mask=None):
        lam_planes = None
        if lam is not None:
            lam = np.asarray(lam, np.complex128)
            lam_planes = np.zeros((2, self.ld))
            lam_planes[0, : self.n], lam_planes[1, : self.n] = lam.real, lam.imag
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = shim().shim_reset(ctypes.byref(self.d), ctypes.byref(self._state()),
                               None if lam_planes is None else lam_planes.ctypes.data,
                               None if m is None else m.ctypes.data,
                               None if self.old_states is None else self.old_states.ctypes.data)
        assert rc == 0
        return self.state()

    def planes_to_obs(self, P):
        M, n = self.M, self.n
        u = (P[0:2 * M:2, :n] + 1j * P[1:2 * M:2, :n]).T
        r = (P[2 * M::2, :n] + 1j * P[2 * M + 1::2, :n]).T
        return np.ascontiguousarray(u), np.ascontiguousarray(r)

    def state(self):
        return self.planes_to_obs(self.S)

    def step(self, actions):
        io = _lib.StepIO()
        if actions is not None:
            a = np.ascontiguousarray(actions)
            self._a = a
            io.action = a.ctypes.data
            w = 2 if np.iscomplexobj(a) else 1
            io.action_env_stride, io.action_comp_stride = a.shape[1] * w, w
        io.reward, io.flags = self.reward.ctypes.data, self.flags.ctypes.data
        io.info_residual, io.info_niter = self.info_res.ctypes.data, self.info_niter.ctypes.data
        io.info_lam, io.terminal_obs = self.info_lam.ctypes.data, self.term.ctypes.data
        io.old_states = None if self.old_states is None else self.old_states.ctypes.data
        if self.entry == "shim_step_certified":
            # trace[i, k] = (r~_k (2M doubles), ||r~_k||, margin_k) of sweep k+1 of env i
            self.trace = np.full((self.n, self.d.max_iters, 2 * self.M + 2), np.nan)
            rc = shim().shim_step_certified(ctypes.byref(self.d), ctypes.byref(self._state()), ctypes.byref(io),
                                            self.trace.ctypes.data, int(getattr(self, "run_fallback", True)))
        elif self.entry == "shim_step_phased":
            stops = np.asarray(self.phase_stops, np.int32)
            rc = shim().shim_step_phased(ctypes.byref(self.d), ctypes.byref(self._state()), ctypes.byref(io),
                                         stops.ctypes.data, len(stops), int(self.phase_split))
        else:
            rc = getattr(shim(), self.entry)(ctypes.byref(self.d), ctypes.byref(self._state()), ctypes.byref(io))
        assert rc == 0
        n = self.n
        f = self.flags[:n]
        u, r = self.state()
        tu, tr = self.planes_to_obs(self.term)
        return dict(u=u, r=r, term_u=tu, term_r=tr, reward=self.reward[:n].copy(), done=(f & 1) != 0, conv=(f & 2) != 0,
                    err=(f & 4) != 0, niter=self.info_niter[:n].copy(), residual=self.info_res[:n].copy(),
                    lam=self.info_lam[:n].copy().view(np.complex128).reshape(n))


def spectral_radius(M, prec_type, lam, qd, *, dt=1.0, Q=None, Qd_fixed=None, grid=None):
    d = _lib.RhoDesc()
    d.M, d.prec_type, d.dt = M, _lib.PREC_TYPES[prec_type], dt
    Q = collocation_matrix(M) if Q is None else Q
    for k, v in enumerate(np.asarray(Q).reshape(-1)):
        d.Q[k] = float(v)
    if Qd_fixed is not None:
        for k, v in enumerate(np.asarray(Qd_fixed).reshape(-1)):
            d.Qd_fixed[k] = float(v)
    qd_arr = None
    if qd is not None:
        qd_arr = np.ascontiguousarray(qd)
        d.qd_is_complex = int(np.iscomplexobj(qd_arr))
        d.qd_broadcast = int(qd_arr.ndim == 1 or qd_arr.shape[0] == 1)
    if grid is not None:
        d.grid_re, d.grid_im, d.re_lo, d.re_hi, d.im_lo, d.im_hi = grid
        N, lam_arr = d.grid_re * d.grid_im, None
    else:
        lam_arr = np.ascontiguousarray(lam, np.complex128).reshape(-1)
        N = lam_arr.shape[0]
    rho = np.zeros(N)
    rc = shim().shim_spectral_radius(ctypes.byref(d), N, None if lam_arr is None else lam_arr.ctypes.data,
                                     None if qd_arr is None else qd_arr.ctypes.data, rho.ctypes.data)
    assert rc == 0
    return rho
