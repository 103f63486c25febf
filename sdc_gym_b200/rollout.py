"""Device-resident rollout collection for the RL caller of the path (``rl_playground.py:286`` ``model.learn`` ->
SB3 ``collect_rollouts``; BASELINE config "PPG rollout collection, sdc-v1, norm_obs").

``RolloutBuffer`` keeps observations (as planes), actions, rewards, episode starts, values and log-probs of
``n_steps`` x ``num_envs`` transitions in HBM and computes returns / GAE advantages with one kernel
(``sdcgym_gae``, SB3 semantics).  ``collect_rollouts`` drives a (normalised) ``SDCVecEnv`` with a policy callable
without any host round trip per step.  The learner itself (PPO/PPG update) is out of scope.
"""
from __future__ import annotations

import ctypes

from . import _lib


def _torch():
    import torch

    return torch


class RolloutBuffer:
    def __init__(self, n_steps, num_envs, obs_planes, action_dim, device, gamma=0.99, gae_lambda=0.95,
                 obs_dtype=None, action_is_complex=False, ld=None):
        torch = _torch()
        self.n_steps, self.num_envs, self.P, self.A = int(n_steps), int(num_envs), int(obs_planes), int(action_dim)
        self.gamma, self.gae_lambda = float(gamma), float(gae_lambda)
        self.device = torch.device(device)
        f64 = torch.float64
        T, N = self.n_steps, self.num_envs
        # observation slots keep the env's plane stride `ld` so that a slot can be the direct output of the
        # normalisation kernel (VecNormalize.step_tensor(obs_out=slot)); `observations` is the (T, P, N) view
        self.ld = N if ld is None else int(ld)
        self.obs_store = torch.zeros((T, self.P, self.ld), dtype=obs_dtype or f64, device=self.device)
        self.observations = self.obs_store[:, :, :N]
        self.actions = torch.zeros((T, N, self.A), dtype=torch.complex128 if action_is_complex else f64,
                                   device=self.device)
        self.rewards = torch.zeros((T, N), dtype=f64, device=self.device)
        self.values = torch.zeros((T, N), dtype=f64, device=self.device)
        self.log_probs = torch.zeros((T, N), dtype=f64, device=self.device)
        self.episode_starts = torch.zeros((T, N), dtype=torch.uint8, device=self.device)
        self.advantages = torch.zeros((T, N), dtype=f64, device=self.device)
        self.returns = torch.zeros((T, N), dtype=f64, device=self.device)
        self.pos = 0
        self.full = False
        self._L = _lib.load()
        self._guard = _lib.DeviceGuard(self.device)

    def reset(self):
        self.pos, self.full = 0, False

    def add(self, obs_planes, actions, rewards, episode_starts, values=None, log_probs=None):
        """Store one transition of every env.  ``obs_planes``: (P, N) tensor (the observation the action was taken
        on), ``episode_starts``: uint8/bool (N,)."""
        if self.pos >= self.n_steps:
            raise IndexError("rollout buffer is full")
        t = self.pos
        self.observations[t].copy_(obs_planes)
        self.actions[t].copy_(actions)
        self.rewards[t].copy_(rewards)
        self.episode_starts[t].copy_(episode_starts)
        if values is not None:
            self.values[t].copy_(values)
        if log_probs is not None:
            self.log_probs[t].copy_(log_probs)
        self.pos += 1
        self.full = self.pos == self.n_steps

    @_lib.on_device
    def compute_returns_and_advantage(self, last_values, dones):
        """SB3 ``RolloutBuffer.compute_returns_and_advantage(last_values, dones)`` on the device."""
        torch = _torch()
        lv = last_values.to(self.device, torch.float64).contiguous()
        ld = dones.to(self.device).to(torch.uint8).contiguous()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._L.sdcgym_gae(self.pos, self.num_envs, self.rewards.data_ptr(), self.values.data_ptr(),
                                      self.episode_starts.data_ptr(), lv.data_ptr(), ld.data_ptr(), self.gamma,
                                      self.gae_lambda, self.advantages.data_ptr(), self.returns.data_ptr(), stream),
                   "sdcgym_gae")
        self._keep = (lv, ld)
        return self.advantages[: self.pos], self.returns[: self.pos]


def collect_rollouts(env, policy, n_steps, buffer=None, gamma=0.99, gae_lambda=0.95):
    """Collect ``n_steps`` transitions of every env on the device.

    ``env``: ``VecNormalize`` (or bare ``SDCVecEnv``) that has been ``reset()``.
    ``policy(obs_planes) -> (actions, values, log_probs)`` with ``obs_planes`` a (4M, N) CUDA tensor; ``values`` /
    ``log_probs`` may be ``None``.  Returns the filled ``RolloutBuffer`` (returns and advantages computed with the
    policy's value of the final observation)."""
    torch = _torch()
    venv = getattr(env, "venv", env)
    N, M = venv.num_envs, venv.M
    normalised = env is not venv and getattr(env, "norm_obs", False)

    if buffer is None:
        buffer = RolloutBuffer(n_steps, N, 4 * M, venv._kernel_n_act or venv.n_act, venv.device, gamma, gae_lambda,
                               action_is_complex=venv.free_action_space, ld=venv.ld)
    buffer.reset()
    # normalised env + a buffer with the env's plane stride: the normalisation kernel writes the next observation
    # straight into its buffer slot (no norm_planes -> buffer copy per step); only slot 0 is filled by a copy
    direct = bool(normalised and buffer.ld == venv.ld and buffer.obs_store.dtype == venv.S.dtype and n_steps > 0)
    if direct:
        buffer.obs_store[0].copy_(env.current_norm_planes)

    def current_obs():
        return env.current_norm_planes[:, :N] if normalised else venv.S[:, :N]

    starts = getattr(env, "_last_episode_starts", None)
    if starts is None:
        starts = torch.ones(N, dtype=torch.uint8, device=venv.device)
    dones = starts
    for _ in range(n_steps):
        t = buffer.pos
        if direct:
            obs = buffer.observations[t]
            actions, values, log_probs = policy(obs)
            nxt = buffer.obs_store[t + 1] if t + 1 < n_steps else env.norm_planes
            out = env.step_tensor(actions if venv._kernel_n_act else None, obs_out=nxt)
        else:
            obs = current_obs()
            actions, values, log_probs = policy(obs)
            buffer.observations[t].copy_(obs)  # before the step overwrites the planes
            out = env.step_tensor(actions if venv._kernel_n_act else None)
        buffer.actions[t].copy_(actions)
        buffer.rewards[t].copy_(out["reward"])
        buffer.episode_starts[t].copy_(starts)
        if values is not None:
            buffer.values[t].copy_(values)
        if log_probs is not None:
            buffer.log_probs[t].copy_(log_probs)
        buffer.pos += 1
        starts = out["flags"] & _lib.FLAG_DONE  # uint8 0/1: the next step's episode_starts
    dones = starts
    buffer.full = buffer.pos == buffer.n_steps
    env._last_episode_starts = starts
    _, last_values, _ = policy(current_obs())
    if last_values is None:
        last_values = torch.zeros(N, dtype=torch.float64, device=venv.device)
    buffer.compute_returns_and_advantage(last_values, dones)
    return buffer


class GraphedRollout:
    """``collect_rollouts`` captured ONCE in a CUDA graph and replayed: one graph launch per rollout instead of
    ``n_steps`` x (policy + four or five kernel launches + a dozen buffer copies) driven from Python.

    Small and mid-size batches - the reference's own regime is 8 envs - are bound by launch and interpreter overhead,
    not by the kernels (a normalised ``sdc-v1`` step of 16 384 envs is < 30 us of GPU work); the graph removes both.
    The replay runs exactly the captured launches on the same buffers, so its results are those of the eager call
    (``tests/test_gpu_normalize_loss.py``).  Restrictions: single rank (the in-kernel peer exchange numbers its rounds
    on the host), a capturable ``policy`` (plain torch ops on its input, default CUDA generator), and the env must
    not be stepped through the host path in between (the graph owns the episode-start flags it carries over).
    """

    def __init__(self, env, policy, n_steps, buffer=None, gamma=0.99, gae_lambda=0.95, warmup=2):
        torch = _torch()
        venv = getattr(env, "venv", env)
        if env is not venv and hasattr(env, "_multi_rank") and env._multi_rank():
            raise _lib.SdcGymError("GraphedRollout: single rank only (the peer exchange is sequenced from the host)")
        self.env, self.venv, self.policy, self.n_steps = env, venv, policy, int(n_steps)
        self.buffer = buffer
        self._args = (gamma, gae_lambda)
        self._stream = torch.cuda.Stream(device=venv.device)
        self._graph = None
        self._warm = max(1, int(warmup))
        self.replays = 0

    def _eager(self):
        g, l = self._args
        self.buffer = collect_rollouts(self.env, self.policy, self.n_steps, buffer=self.buffer, gamma=g, gae_lambda=l)
        return self.buffer

    def collect(self):
        """The next rollout.  The first ``warmup`` calls run eagerly (they configure kernels and warm the allocator),
        the next one is captured while it runs, every later one is a graph replay."""
        torch = _torch()
        with torch.cuda.device(self.venv.device):
            if self._graph is not None:
                self._graph.replay()
                self.replays += 1
                self.venv._step_count += self.n_steps
                self.venv._invalidate()
                self.buffer.pos, self.buffer.full = self.n_steps, True
                return self.buffer
            if self._warm > 0:
                self._warm -= 1
                return self._eager()
            # the episode-start flags a rollout hands to the next one live in ONE static tensor that the graph reads
            # at its first step and rewrites at its end (eagerly, every call leaves a fresh tensor behind)
            carry = self.env._last_episode_starts.clone()
            self.env._last_episode_starts = carry
            cur = torch.cuda.current_stream(self.venv.device)
            self._stream.wait_stream(cur)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(self._stream):
                with torch.cuda.graph(graph, stream=self._stream):
                    self._eager()
                    carry.copy_(self.env._last_episode_starts)
            self.env._last_episode_starts = carry
            cur.wait_stream(self._stream)
            # a capture records, it does not run: replay once so that this call, too, returns a collected rollout
            self._graph = graph
            return self.collect()
