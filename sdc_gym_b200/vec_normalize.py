"""Device-side ``VecNormalize`` for ``SDCVecEnv`` (reference call site ``utils/utils.py:295-312``).

SB3 semantics restated (SB3 is third-party and not installed here - "parity unpinned"):
``RunningMeanStd`` starts at mean 0, var 1, count 1e-4 and merges each batch with Chan's formula;
``obs <- clip((obs - mean) / sqrt(var + eps), +-clip_obs)``; returns ``R <- gamma R + r`` per env (reset on done),
``reward <- clip(r / sqrt(var_R + eps), +-clip_reward)``; ``terminal_observation`` is normalised too.
Deviation (documented): the reference feeds complex128 observations to SB3, whose variance/clip on
complex data is ill-defined; here the statistics are kept per real plane (re and im of every u_m, r_m).

The moments never leave the GPU: accumulate -> (all-reduce over ranks) -> merge -> apply are kernels of
``libsdcgym.so`` (``csrc/vecnorm.cu``); with several ranks the shifted sums are all-reduced so every rank holds
the same normaliser (the shift is the running mean, identical everywhere).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from . import dist as _dist_mod


def _torch():
    import torch

    return torch


class _RunningMeanStd:
    def __init__(self, P, device, epsilon=1e-4):
        torch = _torch()
        self.P = P
        self.mean = torch.zeros(P, dtype=torch.float64, device=device)
        self.var = torch.ones(P, dtype=torch.float64, device=device)
        self.count2 = torch.tensor([epsilon, epsilon], dtype=torch.float64, device=device)

    @property
    def count(self):
        return float(self.count2[0].item())

    def state(self):
        return dict(mean=self.mean.cpu().numpy(), var=self.var.cpu().numpy(), count=self.count)

    def load(self, st):
        torch = _torch()
        self.mean.copy_(torch.as_tensor(st["mean"]))
        self.var.copy_(torch.as_tensor(st["var"]))
        self.count2.fill_(float(st["count"]))


class _StepOut(dict):
    """Result dict of ``VecNormalize.step_tensor``: the normalised terminal planes are produced on first access
    (SB3 normalises ``terminal_observation`` only for the envs that finished; a rollout loop that never looks at
    them should not pay a 2 x 32M-byte pass per env-step for it).  Valid until the next step."""

    def __init__(self, base, terminal_fn):
        super().__init__(base)
        self._terminal_fn = terminal_fn

    def __missing__(self, key):
        if key == "terminal" and self._terminal_fn is not None:
            self["terminal"] = self._terminal_fn()
            return dict.__getitem__(self, "terminal")
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (key == "terminal" and self._terminal_fn is not None)


class VecNormalize:
    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0,
                 gamma=0.99, epsilon=1e-8, sync=True):
        torch = _torch()
        if venv.collect_states:
            raise NotImplementedError("VecNormalize over collect_states buffers is not supported")
        self.venv = venv
        self._L = _lib.load()
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.sync = sync
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        dev, N, M = venv.device, venv.num_envs, venv.M
        self.P = 4 * M
        self.obs_rms = _RunningMeanStd(self.P, dev)
        self.ret_rms = _RunningMeanStd(1, dev)
        self.returns = torch.zeros(max(N, 1), dtype=torch.float64, device=dev)
        self.norm_planes = torch.zeros_like(venv.S)
        self.norm_terminal = torch.zeros_like(venv.S)
        self.current_norm_planes = self.norm_planes  # where the normalised current observation lives (see step_tensor)
        self.norm_reward_buf = torch.zeros(max(N, 1), dtype=torch.float64, device=dev)
        nscr = self._L.sdcgym_vecnorm_scratch_doubles(self.P + 1)  # (+1: the combined obs + returns launch)
        self._scratch = torch.zeros(nscr, dtype=torch.float64, device=dev)
        self._rscratch = torch.zeros(self._L.sdcgym_vecnorm_scratch_doubles(1), dtype=torch.float64, device=dev)
        self.fused_update = True  # single-rank: accumulate + merge in one launch (False: the three-kernel sequence)
        # several ranks: sync=True / "peer" folds the all-reduce of the moment sums into the statistics kernel (P2P
        # stores into every rank's exchange region over NVLink, csrc/vecnorm.cu update_kernel<.., DIST>);
        # sync="nccl" keeps accumulate -> ncclAllReduce -> merge
        self._xchg_obs = self._xchg_ret = self._xchg_both = None
        # shifted sums of the observation planes and of the returns, contiguous so that a multi-rank step needs ONE
        # all-reduce for both statistics
        self._allsums = torch.zeros(2 * self.P + 1 + 3, dtype=torch.float64, device=dev)
        self._sums = self._allsums[: 2 * self.P + 1]
        self._rsums = self._allsums[2 * self.P + 1:]
        self._obs_aos = torch.zeros((max(N, 1), 2, M, 2), dtype=torch.float64, device=dev)
        self.old_reward = None
        self._guard = venv._guard

    # ---- delegation --------------------------------------------------------------------------------
    def __getattr__(self, name):
        return getattr(self.venv, name)

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.venv.device).cuda_stream)

    def _global_count(self, N):
        """number of envs over all ranks (static: reduced once, so no per-step host sync)"""
        if getattr(self, "_global_n", None) is None:
            torch = _torch()
            t = torch.tensor([float(N)], dtype=torch.float64, device=self.venv.device)
            _dist_mod.all_reduce_sum(t)
            self._global_n = float(t.item())
        return self._global_n

    # ---- statistics --------------------------------------------------------------------------------
    def _multi_rank(self):
        return self.sync and _dist_mod.is_distributed()

    def _peer_mode(self):
        """True when the multi-rank statistics go through the in-kernel peer-memory exchange."""
        if not self._multi_rank() or self.sync == "nccl" or not self.fused_update:
            return False
        if self._xchg_obs is None:
            self._xchg_obs = _dist_mod.PeerExchange(2 * self.P + 1)
            self._xchg_ret = _dist_mod.PeerExchange(3)
            self._xchg_both = _dist_mod.PeerExchange(2 * self.P + 3)
        return True

    def _update_both(self, N):
        """Observation planes and discounted returns (advanced in the same pass) in ONE launch - with several ranks
        also one in-kernel exchange for both statistics."""
        v, o, r = self.venv, self.obs_rms, self.ret_rms
        x = ctypes.byref(self._xchg_both.next()) if self._peer_mode() else None
        _lib.check(self._L.sdcgym_vecnorm_update_both(
            self.P, N, v.ld, v.S.data_ptr(), v.reward.data_ptr(), self.gamma, self.returns.data_ptr(),
            o.mean.data_ptr(), o.var.data_ptr(), o.count2.data_ptr(), r.mean.data_ptr(), r.var.data_ptr(),
            r.count2.data_ptr(), self._scratch.data_ptr(), self._sums.data_ptr(), self._rsums.data_ptr(), x,
            self._stream()), "vecnorm_update_both")

    def _update(self, rms, planes_ptr, P, N, ld, sums):
        L, s = self._L, self._stream()
        if self._peer_mode():
            x = self._xchg_obs if P > 1 else self._xchg_ret
            scratch = self._scratch if P > 1 else self._rscratch
            _lib.check(L.sdcgym_vecnorm_update_dist(P, N, ld, planes_ptr, rms.mean.data_ptr(), rms.var.data_ptr(),
                                                    rms.count2.data_ptr(), scratch.data_ptr(), sums.data_ptr(),
                                                    ctypes.byref(x.next()), s), "vecnorm_update_dist")
            return
        if self.fused_update and not self._multi_rank():
            scratch = self._scratch if P > 1 else self._rscratch
            _lib.check(L.sdcgym_vecnorm_update(P, N, ld, planes_ptr, rms.mean.data_ptr(), rms.var.data_ptr(),
                                               rms.count2.data_ptr(), scratch.data_ptr(), sums.data_ptr(), s),
                       "vecnorm_update")
            return
        self._accumulate(rms, planes_ptr, P, N, ld, sums)
        total = float(N)
        if self.sync and _dist_mod.is_distributed():
            _dist_mod.all_reduce_sum(sums)  # < 1 kB, latency bound
            total = self._global_count(N)
        self._merge(rms, P, total, sums)

    def _accumulate(self, rms, planes_ptr, P, N, ld, sums):
        _lib.check(self._L.sdcgym_vecnorm_accumulate(P, N, ld, planes_ptr, rms.mean.data_ptr(), self._scratch.data_ptr(),
                                                     sums.data_ptr(), self._stream()), "vecnorm_accumulate")

    def _merge(self, rms, P, total, sums):
        _lib.check(self._L.sdcgym_vecnorm_merge(P, total, sums.data_ptr(), rms.mean.data_ptr(), rms.var.data_ptr(),
                                                rms.count2.data_ptr(), self._stream()), "vecnorm_merge")

    def _normalize_planes(self, src, dst):
        v = self.venv
        _lib.check(self._L.sdcgym_vecnorm_apply(self.P, v.num_envs, v.ld, src.data_ptr(), self.obs_rms.mean.data_ptr(),
                                                self.obs_rms.var.data_ptr(), self.epsilon, self.clip_obs,
                                                dst.data_ptr(), self._stream()), "vecnorm_apply")

    def _obs_out(self, planes):
        """planes (4M, ld) -> (N, 2, M) complex in the env's output mode"""
        torch = _torch()
        v = self.venv
        _lib.check(self._L.sdcgym_export_obs(v.M, v.num_envs, v.ld, planes.data_ptr(), self._obs_aos.data_ptr(),
                                             self._stream()), "export_obs")
        t = torch.view_as_complex(self._obs_aos)[: v.num_envs]
        return t if v.output == "torch" else t.cpu().numpy()

    # ---- VecEnv API --------------------------------------------------------------------------------
    @_lib.on_device
    def reset(self, **kw):
        v = self.venv
        v.reset(**kw)
        self.returns.zero_()
        if self.norm_obs:
            if self.training:
                self._update(self.obs_rms, v.S.data_ptr(), self.P, v.num_envs, v.ld, self._sums)
            self._normalize_planes(v.S, self.norm_planes)
            self.current_norm_planes = self.norm_planes
            return self._obs_out(self.norm_planes)
        return self._obs_out(v.S)

    def _normalized_terminal(self):
        self._normalize_planes(self.venv.terminal, self.norm_terminal)
        return self.norm_terminal[:, : self.venv.num_envs]

    @_lib.on_device
    def step_tensor(self, actions=None, obs_out=None):
        """Device-resident normalised step.  Returns the env's dict plus ``obs_planes`` (normalised S planes),
        ``reward`` replaced by the normalised reward and ``raw_reward``; ``terminal`` (normalised terminal planes)
        is computed when first read.  ``obs_out``: optional (4M, ld) float64 CUDA tensor that receives the
        normalised observation instead of ``self.norm_planes`` (a rollout buffer slot: saves one pass)."""
        v, L, s = self.venv, self._L, self._stream()
        raw = v.step_tensor(actions)
        N = v.num_envs
        # several ranks, both statistics live: accumulate both, ONE all-reduce (the only collective of a normalised
        # step), merge both - instead of one collective per statistic
        peer = self.training and self._peer_mode()
        combined = self.training and self.norm_obs and self._multi_rank() and not peer
        # single rank or in-kernel exchange, both statistics live: ONE launch for both
        both = self.training and self.norm_obs and self.fused_update and (peer or not self._multi_rank())
        if both:
            self._update_both(N)
        if combined:
            self._accumulate(self.obs_rms, v.S.data_ptr(), self.P, N, v.ld, self._sums)
            _lib.check(L.sdcgym_vecnorm_returns(N, v.reward.data_ptr(), self.gamma, self.returns.data_ptr(), s), "returns")
            self._accumulate(self.ret_rms, self.returns.data_ptr(), 1, N, max(N, 1), self._rsums)
            _dist_mod.all_reduce_sum(self._allsums)
            total = self._global_count(N)
            self._merge(self.obs_rms, self.P, total, self._sums)
            self._merge(self.ret_rms, 1, total, self._rsums)
        if self.norm_obs:
            out = _StepOut(raw, self._normalized_terminal if self.venv.terminal is not None else None)
            out.pop("terminal", None)
            if self.training and not combined and not both:
                self._update(self.obs_rms, v.S.data_ptr(), self.P, N, v.ld, self._sums)
            dst = self.norm_planes
            if obs_out is not None:
                if (obs_out.shape != v.S.shape or obs_out.dtype != v.S.dtype or not obs_out.is_contiguous()
                        or obs_out.device != v.S.device):
                    raise ValueError(f"obs_out must be a contiguous {tuple(v.S.shape)} float64 tensor on {v.S.device}")
                dst = obs_out
            self._normalize_planes(v.S, dst)
            self.current_norm_planes = dst
            out["obs_planes"] = dst[:, :N]
        else:
            out = dict(raw)
            out["obs_planes"] = v.S[:, :N]
        out["raw_reward"] = raw["reward"]
        if combined or both:
            pass  # return statistics already updated above
        elif peer:
            r = self.ret_rms
            _lib.check(L.sdcgym_vecnorm_update_returns_dist(N, v.reward.data_ptr(), self.gamma, self.returns.data_ptr(),
                                                            r.mean.data_ptr(), r.var.data_ptr(), r.count2.data_ptr(),
                                                            self._rscratch.data_ptr(), self._rsums.data_ptr(),
                                                            ctypes.byref(self._xchg_ret.next()), s),
                       "vecnorm_update_returns_dist")
        elif self.training and self.fused_update and not self._multi_rank():
            r = self.ret_rms
            _lib.check(L.sdcgym_vecnorm_update_returns(N, v.reward.data_ptr(), self.gamma, self.returns.data_ptr(),
                                                       r.mean.data_ptr(), r.var.data_ptr(), r.count2.data_ptr(),
                                                       self._rscratch.data_ptr(), self._rsums.data_ptr(), s),
                       "vecnorm_update_returns")
        elif self.training:
            _lib.check(L.sdcgym_vecnorm_returns(N, v.reward.data_ptr(), self.gamma, self.returns.data_ptr(), s), "returns")
            self._update(self.ret_rms, self.returns.data_ptr(), 1, N, max(N, 1), self._rsums)
        _lib.check(L.sdcgym_vecnorm_reward(N, v.reward.data_ptr(), v.flags.data_ptr(), self.ret_rms.var.data_ptr(),
                                           self.epsilon, self.clip_reward, int(self.norm_reward),
                                           self.norm_reward_buf.data_ptr(), self.returns.data_ptr(), s), "reward")
        out["reward"] = self.norm_reward_buf[:N]
        return out

    def _vecnorm_struct(self):
        vn = getattr(self, "_vn_struct", None)
        if vn is None:
            vn = _lib.VecNorm()
            o, r = self.obs_rms, self.ret_rms
            vn.obs_mean, vn.obs_var, vn.obs_count2 = o.mean.data_ptr(), o.var.data_ptr(), o.count2.data_ptr()
            vn.ret_mean, vn.ret_var, vn.ret_count2 = r.mean.data_ptr(), r.var.data_ptr(), r.count2.data_ptr()
            vn.returns = self.returns.data_ptr()
            vn.scratch_obs, vn.scratch_ret = self._scratch.data_ptr(), self._rscratch.data_ptr()
            vn.sums_obs, vn.sums_ret = self._sums.data_ptr(), self._rsums.data_ptr()
            vn.out_planes, vn.out_reward = self.norm_planes.data_ptr(), self.norm_reward_buf.data_ptr()
            self._vn_struct = vn
        vn.norm_obs, vn.norm_reward, vn.training = int(self.norm_obs), int(self.norm_reward), int(self.training)
        vn.gamma, vn.epsilon = float(self.gamma), float(self.epsilon)
        vn.clip_obs, vn.clip_reward = float(self.clip_obs), float(self.clip_reward)
        return vn

    def _native_host_step(self, actions):
        """numpy in / numpy out through ONE C call (``sdcgym_pipe_step_block`` with the normaliser attached): upload,
        step, statistics, normalisation, one download - the reference's training regime (8 envs behind VecNormalize)
        is pure latency, and the call-by-call path costs ~200 us per step against ~50 us here."""
        v = self.venv
        obs, rewards, dones, infos = v._step_host(actions, vn=self._vecnorm_struct())
        if self.norm_obs:
            self.current_norm_planes = self.norm_planes
        self.old_reward = v.reward[: v.num_envs]
        stamp = v._step_count

        def fetch():
            if v._step_count != stamp:
                raise RuntimeError("terminal observations of this step are gone: the env has stepped since")
            return self._terminal_host(self._terminal_planes_full())

        infos._terminal_fetch = fetch
        return obs, rewards, dones, infos

    @_lib.on_device
    def step(self, actions):
        """(obs, rewards, dones, infos) with normalised obs / rewards (numpy or torch per the env's ``output``)."""
        torch = _torch()
        v = self.venv
        a = actions
        if (v.output != "torch" and not (isinstance(a, torch.Tensor) and a.is_cuda) and v.host_pipeline == "native"
                and self.fused_update and not self._multi_rank() and not v._rho_reward):
            return self._native_host_step(a)
        if v._kernel_n_act > 0 and not (isinstance(a, torch.Tensor) and a.is_cuda):
            arr = np.asarray(a, dtype=np.complex128 if v.free_action_space else np.float64).reshape(v.num_envs, -1)
            a = torch.as_tensor(arr).to(v.device)
        out = self.step_tensor(a if v._kernel_n_act else None)
        N = v.num_envs
        obs = self._obs_out(out["obs_planes"] if self.norm_obs else v.S)
        self.old_reward = out["raw_reward"]
        if v.output == "torch":
            return obs, out["reward"], (out["flags"] & 1).bool(), out
        from .vec_env import MAX_EPISODE_STEPS, LazyInfos, _TruncatedKey

        flags = out["flags"].cpu().numpy()
        dones = (flags & 1).astype(bool)
        niter = out["niter"].cpu().numpy()
        lam = out["lam"].cpu().numpy()
        trunc = _TruncatedKey(niter, MAX_EPISODE_STEPS[v.envname])
        infos = LazyInfos(niter, out["residual"].cpu().numpy(), lam, dones, trunc,
                          lambda: self._terminal_host(self._terminal_planes_full()))
        infos.flags = flags
        return obs, out["reward"].cpu().numpy(), dones, infos

    def _terminal_planes_full(self):
        if self.venv.terminal is None:
            raise RuntimeError("terminal observations are not kept (keep_terminal=False)")
        if not self.norm_obs:
            return self.venv.terminal
        self._normalize_planes(self.venv.terminal, self.norm_terminal)
        return self.norm_terminal

    def _terminal_host(self, planes):
        t = self._obs_out(planes)
        return t.cpu().numpy() if hasattr(t, "cpu") else t

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        return self.step(self._pending)

    def get_original_reward(self):
        return None if self.old_reward is None else self.old_reward.cpu().numpy()

    @_lib.on_device
    def get_original_obs(self):
        return self._obs_out(self.venv.S)

    def normalize_obs(self, obs):
        """Normalise a host observation array (..., 2, M) complex128 with the current statistics."""
        mean, var = self.obs_rms.mean.cpu().numpy(), self.obs_rms.var.cpu().numpy()
        x = np.asarray(obs, dtype=np.complex128)
        flat = x.reshape(-1, x.shape[-2] * x.shape[-1]).view(np.float64)
        y = np.clip((flat - mean) / np.sqrt(var + self.epsilon), -self.clip_obs, self.clip_obs)
        return y.view(np.complex128).reshape(x.shape)

    # ---- persistence (SB3: VecNormalize.save / load, utils/utils.py:297,415-417) -----------------------------
    # The reference's `--env_path` files are SB3 pickles of the wrapper object (complex (2, M) obs_rms).  They are
    # not readable here (SB3 is not a dependency, unpickling runs arbitrary code, and the statistics kept here are
    # per REAL plane - 4M entries - see the module docstring), so `load` rejects them with a clear message.  The own
    # format is a versioned .npz without pickled objects.
    FILE_FORMAT = "sdc_gym_b200.VecNormalize"
    FILE_VERSION = 1

    def state_dict(self):
        return dict(obs_rms=self.obs_rms.state(), ret_rms=self.ret_rms.state(), returns=self.returns.cpu().numpy(),
                    clip_obs=self.clip_obs, clip_reward=self.clip_reward, gamma=self.gamma, epsilon=self.epsilon,
                    norm_obs=self.norm_obs, norm_reward=self.norm_reward, training=self.training)

    def load_state_dict(self, sd, restore_flags=False):
        """statistics, returns and hyper-parameters; ``restore_flags`` also restores norm_obs / norm_reward / training
        (what ``VecNormalize.load`` does - a file resumes the wrapper as it was saved; an explicit
        ``load_state_dict`` into an object the caller configured keeps the caller's flags)"""
        torch = _torch()
        if not isinstance(sd, dict) or "obs_rms" not in sd or "ret_rms" not in sd:
            raise TypeError("not a sdc_gym_b200 VecNormalize state dict (an SB3 VecNormalize object cannot be loaded: "
                            "its statistics are complex (2, M), the device normaliser keeps 4M real planes)")
        if np.asarray(sd["obs_rms"]["mean"]).shape != (self.P,):
            raise ValueError(f"normaliser statistics are for {np.asarray(sd['obs_rms']['mean']).shape[0] // 4} nodes, "
                             f"this env has M={self.P // 4}")
        self.obs_rms.load(sd["obs_rms"])
        self.ret_rms.load(sd["ret_rms"])
        ret = torch.as_tensor(np.asarray(sd["returns"], dtype=np.float64))
        if ret.numel() == self.returns.numel():  # per-env running returns only make sense for the same batch
            self.returns.copy_(ret)
        else:
            self.returns.zero_()
        for k in ("clip_obs", "clip_reward", "gamma", "epsilon"):
            setattr(self, k, float(sd[k]))
        if restore_flags:
            for k in ("norm_obs", "norm_reward", "training"):
                if k in sd:
                    setattr(self, k, bool(sd[k]))

    def save(self, path):
        sd = self.state_dict()
        with open(path, "wb") as f:
            np.savez(f, format=np.array(self.FILE_FORMAT), version=np.array(self.FILE_VERSION),
                     obs_mean=sd["obs_rms"]["mean"], obs_var=sd["obs_rms"]["var"], obs_count=np.array(sd["obs_rms"]["count"]),
                     ret_mean=sd["ret_rms"]["mean"], ret_var=sd["ret_rms"]["var"], ret_count=np.array(sd["ret_rms"]["count"]),
                     returns=sd["returns"],
                     **{k: np.array(sd[k]) for k in ("clip_obs", "clip_reward", "gamma", "epsilon", "norm_obs",
                                                     "norm_reward", "training")})

    @staticmethod
    def load(path, venv):
        with open(path, "rb") as f:
            magic = f.read(2)
        if magic != b"PK":  # not a zip container: most likely an SB3 `VecNormalize.save` pickle (utils/utils.py:297)
            raise ValueError(f"{path}: not a sdc_gym_b200 VecNormalize file (.npz).  Pickles written by stable-baselines' "
                             "VecNormalize.save are not supported: re-create the statistics with this wrapper "
                             "(e.g. a dry run, rl_playground.py:58-82) and save them with VecNormalize.save")
        with np.load(path, allow_pickle=False) as z:
            if "format" not in z.files or str(z["format"]) != VecNormalize.FILE_FORMAT:
                raise ValueError(f"{path}: not a sdc_gym_b200 VecNormalize file")
            if int(z["version"]) > VecNormalize.FILE_VERSION:
                raise ValueError(f"{path}: written by a newer version ({int(z['version'])})")
            sd = dict(obs_rms=dict(mean=z["obs_mean"], var=z["obs_var"], count=float(z["obs_count"])),
                      ret_rms=dict(mean=z["ret_mean"], var=z["ret_var"], count=float(z["ret_count"])),
                      returns=z["returns"],
                      **{k: z[k].item() for k in ("clip_obs", "clip_reward", "gamma", "epsilon", "norm_obs",
                                                  "norm_reward", "training")})
        vn = VecNormalize(venv)
        vn.load_state_dict(sd, restore_flags=True)
        return vn

    def close(self):
        self.venv.close()


class VecCheckNan:
    """``VecCheckNan(env, raise_exception=True)`` as used by the reference under ``--debug_nans``
    (``utils/utils.py:313-314``): raises ``ValueError`` when an action, observation or reward is NaN / Inf."""

    def __init__(self, venv, raise_exception=True, warn_once=True, check_inf=True):
        self.venv, self.raise_exception, self.check_inf = venv, raise_exception, check_inf

    def __getattr__(self, name):
        return getattr(self.venv, name)

    def _check(self, **arrays):
        for name, x in arrays.items():
            if x is None:
                continue
            if hasattr(x, "is_cuda"):
                t = _torch().view_as_real(x) if x.is_complex() else x
                bad = bool((t.isnan() | (t.isinf() if self.check_inf else False)).any())
            else:
                a = np.asarray(x)
                if a.dtype == object:
                    continue
                bad = bool(np.isnan(a).any() or (self.check_inf and np.isinf(a).any()))
            if bad:
                msg = f"found nan or inf in {name}"
                if self.raise_exception:
                    raise ValueError(msg)
                import warnings
                warnings.warn(msg)

    def reset(self, **kw):
        obs = self.venv.reset(**kw)
        self._check(observations=obs)
        return obs

    def step(self, actions):
        if self.venv.prec is None:
            self._check(actions=actions)
        obs, rew, done, infos = self.venv.step(actions)
        self._check(observations=obs, rewards=rew)
        return obs, rew, done, infos

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        return self.step(self._pending)
