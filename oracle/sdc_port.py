"""TEST / BASELINE INFRASTRUCTURE ONLY - numpy restatement ("port") of the reference env's CPU path.

``PortFullEnv`` / ``PortStepEnv`` restate ``SDC_Full_Env`` / ``SDC_Step_Env`` (``sdc_gym/envs/sdc_env.py:15-572``)
with the *same numpy calls in the same order* (``np.linalg.inv``, ``@``, ``np.linalg.norm(., inf)``,
``np.interp``, ``math.log``), one env at a time, so that

* on the same host it is bit-identical to the reference by construction (checked against the golden vectors
  in ``tests/test_port_golden.py`` and live against the reference in the build container), and
* its speed is the reference's speed: this is what ``bench.py`` times as ``cpu_baseline`` / ``--impl reference``
  on the GPU box, where ``/root/reference`` (pure Python, needs gym/pySDC) cannot travel.

gym / pySDC / matplotlib are not needed: the collocation matrix comes from ``sdc_gym_b200.collocation`` (the
same bits the product uses) and lambda is drawn from a ``numpy.random.RandomState(seed)`` like old gym did.
Nothing under ``sdc_gym_b200/`` imports this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg

from sdc_gym_b200.collocation import CollGaussRadauRight

_MIN = {
    7: [0.15223871397682717, 0.12625448001038536, 0.08210714764924298, 0.03994434742760019,
        0.1052662547386142, 0.14075805578834127, 0.15636085758812895],
    5: [0.2818591930905709, 0.2011358490453793, 0.06274536689514164, 0.11790265267514095, 0.1571629578515223],
    4: [0.3198786751412953, 0.08887606314792469, 0.1812366328324738, 0.23273925017954],
    3: [0.3203856825077055, 0.1399680686269595, 0.3716708461097372],
}


class PortFullEnv:
    """sdc-v0 (sdc_env.py:15-497)."""

    max_iters = 50

    def __init__(self, M=None, dt=None, restol=None, prec=None, seed=None, lambda_real_interval=(-100, 0),
                 lambda_imag_interval=(0, 0), lambda_real_interpolation_interval=None, norm_factor=1,
                 residual_weight=0.5, step_penalty=0.1, reward_iteration_only=None,
                 reward_strategy="iteration_only", collect_states=False, use_doubles=True, do_scale=True,
                 free_action_space=False, prec_type="diag"):
        self.M, self.dt, self.restol, self.prec = M, dt, restol, prec
        self.coll = CollGaussRadauRight(M, 0, 1)
        self.Q = self.coll.Qmat[1:, 1:]
        self.u0 = np.ones(M, dtype=np.complex128)
        self.lambda_real_interval = list(lambda_real_interval)
        self.lambda_real_interval_reversed = list(reversed(self.lambda_real_interval))
        self.lambda_imag_interval = list(lambda_imag_interval)
        self.lambda_real_interpolation_interval = lambda_real_interpolation_interval
        self.norm_factor, self.residual_weight, self.step_penalty = norm_factor, residual_weight, step_penalty
        if reward_iteration_only is None:
            self.reward_strategy = reward_strategy.lower()
        elif reward_iteration_only:
            self.reward_strategy = "iteration_only"
        else:
            self.reward_strategy = "residual_change"
        self.collect_states, self.do_scale = collect_states, do_scale
        if free_action_space:  # sdc_env.py:95-110
            self.action_dtype = np.complex128 if use_doubles else np.complex64
        else:
            self.action_dtype = np.float64 if use_doubles else np.float32
        self.prec_type = prec_type
        self.num_episodes = 0
        self.np_random = np.random.RandomState(seed)
        self.state = None
        self.niter = None
        self.lam = None
        self.C = None
        self.initial_residual = None
        if collect_states:
            self.old_states = np.zeros((M * 2, self.max_iters), dtype=np.complex128)

    # -- sdc_env.py:125-132
    def _scale_action(self, action):
        return np.interp(action, (-1, 1), (0, 1)) if self.do_scale else action

    # -- sdc_env.py:134-191 (+ dp_playground.py:194-207 layouts for the learned non-diagonal types)
    def _get_prec(self, scaled_action):
        M = self.M
        if self.prec is None:
            if self.prec_type == "diag":
                Qdmat = np.zeros_like(self.Q, dtype=self.action_dtype)
                np.fill_diagonal(Qdmat, scaled_action)
            elif self.prec_type == "lower_diag":
                Qdmat = np.diag(np.asarray(scaled_action).astype(self.action_dtype), k=-1)
            else:
                out = np.asarray(scaled_action)
                Qdmat = np.zeros((M, M), dtype=self.action_dtype)
                idx = np.tril_indices(M) if self.prec_type == "lower_tri" else np.tril_indices(M, k=-1)
                Qdmat[idx] = out
        elif self.prec.upper() == "LU":
            _, _, U = scipy.linalg.lu(np.array(self.Q.T, order="C", copy=True))
            Qdmat = U.T
        elif self.prec.lower() == "min":
            Qdmat = np.zeros_like(self.Q)
            np.fill_diagonal(Qdmat, _MIN.get(M, np.zeros(M)))
        elif self.prec.upper() == "EE":
            Qdmat = np.zeros_like(self.Q)
            for m in range(M):
                Qdmat[m, 0:m] = self.coll.delta_m[1:m + 1]
        elif self.prec.lower() == "zeros":
            Qdmat = np.zeros_like(self.Q)
        else:
            raise NotImplementedError()
        return Qdmat

    # -- sdc_env.py:193-207
    def _compute_pinv(self, scaled_action):
        Qdmat = self._get_prec(scaled_action=scaled_action)
        return np.linalg.inv(np.eye(self.M) - self.lam * self.dt * Qdmat)

    def _compute_residual(self, u):
        return self.u0 - self.C @ u

    def _inf_norm(self, v):
        return np.linalg.norm(v, np.inf)

    # -- sdc_env.py:209-273
    def step(self, action):
        u, old_residual = self.state
        scaled_action = self._scale_action(action)
        Pinv = self._compute_pinv(scaled_action)
        norm_res_old = self._inf_norm(old_residual)
        residual = old_residual
        done = False
        err = False
        self.niter = 0
        while not done and not self.niter >= self.max_iters:
            self.niter += 1
            u += Pinv @ residual
            residual = self._compute_residual(u)
            norm_res = self._inf_norm(residual)
            err = np.isnan(norm_res) or np.isinf(norm_res)
            if self.collect_states and self.niter < self.max_iters:
                self.old_states[:, self.niter] = np.concatenate((u, residual))
            err = err or norm_res > norm_res_old * 100
            if err:
                reward = -self.step_penalty * (self.max_iters + 1)
                break
            done = norm_res < self.restol
        if not err:
            reward = self.reward_func(self.initial_residual, residual, done, self.niter, scaled_action, Pinv)
        done = True
        self.state = (u, residual)
        info = {"residual": norm_res, "niter": self.niter, "lam": self.lam}
        return (self.old_states if self.collect_states else self.state, reward, done, info)

    # -- sdc_env.py:275-332
    def _generate_lambda(self):
        if self.lambda_real_interpolation_interval is not None:
            lam_low = np.interp(self.num_episodes, self.lambda_real_interpolation_interval,
                                self.lambda_real_interval_reversed)
        else:
            lam_low = self.lambda_real_interval[0]
        self.lam = (1 * self.np_random.uniform(low=lam_low, high=self.lambda_real_interval[1])
                    + 1j * self.np_random.uniform(low=self.lambda_imag_interval[0], high=self.lambda_imag_interval[1]))

    def set_lambda(self, lam):
        """Inject lambda into a freshly reset episode (the statements of _compute_initial_state, :302-314)."""
        self.lam = complex(lam)
        self.C = np.eye(self.M) - self.lam * self.dt * self.Q
        u = np.ones(self.M, dtype=np.complex128)
        residual = self._compute_residual(u)
        self.initial_residual = residual
        self.state = (u, residual)
        if self.collect_states:
            self.old_states[:, 0] = np.concatenate(self.state)
            self.old_states[:, 1:] = 0
        return self.state

    def reset(self):
        self.num_episodes += 1
        self.niter = 0
        self._generate_lambda()
        self.set_lambda(self.lam)
        return self.old_states if self.collect_states else self.state

    # -- sdc_env.py:334-463
    def reward_func(self, old_residual, residual, reached_convergence, steps, scaled_action, Pinv):
        s = self.reward_strategy
        if s == "iteration_only":
            return -steps * self.step_penalty
        if s == "residual_change":
            reward = abs((math.log(self._inf_norm(old_residual * self.norm_factor))
                          - math.log(self._inf_norm(residual * self.norm_factor)))
                         / (math.log(self._inf_norm(self.initial_residual * self.norm_factor))
                            - math.log(self.restol * self.norm_factor)))
            reward *= self.residual_weight
            reward -= steps * self.step_penalty
            return reward
        norm_res = self._inf_norm(residual)
        extra_fact = (self.max_iters + 1 - steps) ** 2 * 10 if reached_convergence else 1
        if s == "gauss_kernel":
            return (1 * np.exp(-(norm_res * (1 / self.restol)) ** 2 / 2)) * extra_fact
        if s in ("fast_convergence", "smooth_fast_convergence", "smoother_fast_convergence"):
            reward = 1000 if norm_res == 0 else -math.log(norm_res)
            if s == "smooth_fast_convergence" and reward > 1:
                reward = 1 + math.log(reward)
            reward *= extra_fact
            if s == "smoother_fast_convergence" and reward > 1:
                reward = 1 + math.log(reward)
            return reward
        if s == "spectral_radius":
            Qdmat = self._get_prec(scaled_action)
            mulpinv = Pinv.dot(self.Q - Qdmat)
            return max(abs(np.linalg.eigvals(self.lam * self.dt * mulpinv)))
        raise NotImplementedError(s)


class PortStepEnv(PortFullEnv):
    """sdc-v1 (sdc_env.py:499-572)."""

    def step(self, action):
        u, old_residual = self.state
        scaled_action = self._scale_action(action)
        Pinv = self._compute_pinv(scaled_action)
        u += Pinv @ old_residual
        residual = self._compute_residual(u)
        norm_res = self._inf_norm(residual)
        norm_res_old = self._inf_norm(old_residual)
        self.niter += 1
        err = np.isnan(norm_res) or np.isinf(norm_res)
        err = err or norm_res > norm_res_old * 100
        done = norm_res < self.restol
        if not err:
            reward = self.reward_func(old_residual, residual, done, self.niter, scaled_action, Pinv)
        else:
            reward = -self.step_penalty * (self.max_iters + 1)
        done = done or self.niter >= self.max_iters or err
        self.state = (u, residual)
        if self.collect_states and self.niter < self.max_iters:
            self.old_states[:, self.niter] = np.concatenate(self.state)
        info = {"residual": norm_res, "niter": self.niter, "lam": self.lam}
        return (self.old_states if self.collect_states else self.state, reward, done, info)


ENV_CLASSES = {"sdc-v0": PortFullEnv, "sdc-v1": PortStepEnv}


class PortDummyVecEnv:
    """The DummyVecEnv + TimeLimit loop the reference wraps its envs in (utils/utils.py:284-294), restated:
    step envs one by one; on done store ``terminal_observation`` and reset.  ``parity unpinned`` for SB3 itself
    (not installed); this follows its documented semantics."""

    def __init__(self, kind, num_envs, seed=None, **kwargs):
        self.kind = kind
        self.envs = [ENV_CLASSES[kind](seed=None if seed is None else seed + i, **kwargs) for i in range(num_envs)]
        self.num_envs = num_envs
        self.max_episode_steps = {"sdc-v0": 1, "sdc-v1": 50}[kind]
        self._elapsed = [0] * num_envs

    def reset(self):
        self._elapsed = [0] * self.num_envs
        return np.stack([np.stack(e.reset()) for e in self.envs])

    def step(self, actions):
        obs, rews, dones, infos = [], [], [], []
        for i, (e, a) in enumerate(zip(self.envs, actions)):
            o, r, d, info = e.step(a)
            self._elapsed[i] += 1
            if self._elapsed[i] >= self.max_episode_steps:
                info["TimeLimit.truncated"] = not d
                d = True
            if d:
                info["terminal_observation"] = np.stack(o).copy()
                o = e.reset()
                self._elapsed[i] = 0
            obs.append(np.stack(o).copy())
            rews.append(r)
            dones.append(d)
            infos.append(info)
        return np.stack(obs), np.array(rews), np.array(dones), infos


def rollout_throughput(kind, seconds, *, num_envs=8, M=5, seed=0, **kwargs):
    """Random-action rollout for about ``seconds`` of wall clock; returns (env_steps, elapsed, sum_niter)."""
    import time

    kw = dict(M=M, dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
    kw.update(kwargs)
    vec = PortDummyVecEnv(kind, num_envs, seed=seed, **kw)
    rng = np.random.RandomState(seed + 12345)
    vec.reset()
    steps = 0
    sum_niter = 0
    t0 = time.perf_counter()
    while True:
        actions = rng.uniform(-1, 1, (num_envs, M))
        _, _, _, infos = vec.step(list(actions))
        steps += num_envs
        sum_niter += sum(i["niter"] for i in infos) if kind == "sdc-v0" else num_envs
        el = time.perf_counter() - t0
        if el >= seconds:
            return steps, el, sum_niter
