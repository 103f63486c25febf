"""Batched forward values of the ``dp_playground.py`` losses on the device.

``SpectralRadiusLoss(M, dt, prec_type)(lams, outputs)`` mirrors ``dp_playground.py:186-231``: the mean over the
batch of ``rho(lam*dt * inv(I - lam*dt*Qd) @ (Q - Qd))`` with ``Qd = get_qdmat(output)``.  The reference runs
``jnp.linalg.inv`` + ``jnp.linalg.eigvals`` per sample on the CPU (it forces JAX onto the CPU because eigvals
has no GPU kernel, ``dp_playground.py:981-985``); here one thread per sample runs a complex Hessenberg-QR
(``csrc/specrad.cuh``) - 1e-10 relative agreement with LAPACK is the contract, not bit equality.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .collocation import collocation_matrix
from .precond import PREC_TYPES, fixed_preconditioner, num_actions


class UnknownPrecTypeError(ValueError):
    pass


def _torch():
    import torch

    return torch


class SpectralRadiusLoss:
    def __init__(self, M, dt, prec_type="diag", *, prec=None, Q=None, device=None):
        if prec is None and prec_type not in PREC_TYPES:
            raise UnknownPrecTypeError(prec_type)
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.SdcGymError("SpectralRadiusLoss needs a CUDA device: there is no CPU fallback")
        self._L = _lib.load()
        self.M, self.dt = int(M), float(dt)
        self.prec = prec
        self.prec_type = "fixed" if prec is not None else prec_type
        self.Q = collocation_matrix(self.M) if Q is None else np.ascontiguousarray(Q, dtype=np.float64)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_out = 0 if prec is not None else num_actions(self.M, prec_type)
        d = _lib.RhoDesc()
        d.M, d.prec_type, d.dt = self.M, _lib.PREC_TYPES[self.prec_type], self.dt
        for k, v in enumerate(self.Q.reshape(-1)):
            d.Q[k] = float(v)
        if prec is not None:
            for k, v in enumerate(fixed_preconditioner(prec, self.M, self.Q).reshape(-1)):
                d.Qd_fixed[k] = float(v)
        self._desc = d
        self._guard = _lib.DeviceGuard(self.device)

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _outputs_tensor(self, outputs, B):
        torch = _torch()
        if self.prec is not None:
            return None, 0, 0
        t = outputs if isinstance(outputs, torch.Tensor) else torch.as_tensor(np.array(outputs))
        t = t.to(self.device)
        is_c = t.is_complex()
        t = t.to(torch.complex128 if is_c else torch.float64)
        if t.dim() == 1:
            t = t.reshape(1, -1)
        if t.shape[-1] != self.n_out:
            raise ValueError(f"outputs must have {self.n_out} components for prec_type={self.prec_type}")
        broadcast = int(t.shape[0] == 1 and B != 1)
        if not broadcast and t.shape[0] != B:
            raise ValueError("outputs and lams disagree on the batch size")
        t = t.contiguous()
        return (torch.view_as_real(t).contiguous() if is_c else t), int(is_c), broadcast

    @_lib.on_device
    def spectral_radii(self, lams, outputs=None):
        """rho per sample as a CUDA float64 tensor (B,).  ``lams``: (B,) or (B, 1) complex (numpy or torch)."""
        torch = _torch()
        lam = lams if isinstance(lams, torch.Tensor) else torch.as_tensor(np.asarray(lams, dtype=np.complex128))
        lam = lam.to(self.device).to(torch.complex128).reshape(-1).contiguous()
        B = lam.numel()
        if B == 0:  # an empty shard: nothing to launch (a NULL lambda pointer would select grid mode in the ABI)
            return torch.empty(0, dtype=torch.float64, device=self.device)
        qd, is_c, bc = self._outputs_tensor(outputs, B)
        d = self._desc
        d.qd_is_complex, d.qd_broadcast = is_c, bc
        d.grid_re = d.grid_im = 0
        rho = torch.empty(B, dtype=torch.float64, device=self.device)
        lam_r = torch.view_as_real(lam)
        _lib.check(self._L.sdcgym_spectral_radius(ctypes.byref(d), B, lam_r.data_ptr(),
                                                  None if qd is None else qd.data_ptr(), rho.data_ptr(),
                                                  self._stream()), "sdcgym_spectral_radius")
        self._keep = (lam_r, qd)
        return rho

    def __call__(self, lams, outputs=None):
        """mean spectral radius over the batch (``jnp.mean(jax.vmap(...))``, dp_playground.py:230-231)."""
        rho = self.spectral_radii(lams, outputs)
        return self.mean(rho)

    @_lib.on_device
    def radii_and_grads(self, lams, outputs):
        """(rho (B,), g (B, n_out) complex128) with d rho_b = Re(sum_k g_bk d output_bk): per-sample spectral
        radius and its derivative with respect to the Q_delta parameters (left/right eigenvector formula,
        ``csrc/specrad.cuh``).  Valid where the dominant eigenvalue is simple."""
        torch = _torch()
        if self.prec is not None:
            raise ValueError("a fixed preconditioner has no parameters to differentiate")
        lam = lams if isinstance(lams, torch.Tensor) else torch.as_tensor(np.asarray(lams, dtype=np.complex128))
        lam = lam.detach().to(self.device).to(torch.complex128).reshape(-1).contiguous()
        B = lam.numel()
        if B == 0:
            return (torch.empty(0, dtype=torch.float64, device=self.device),
                    torch.empty((0, self.n_out), dtype=torch.complex128, device=self.device))
        out = outputs.detach() if isinstance(outputs, torch.Tensor) else outputs
        qd, is_c, bc = self._outputs_tensor(out, B)
        if bc:
            qd = qd.expand(B, *qd.shape[1:]).contiguous()
        d = self._desc
        d.qd_is_complex, d.qd_broadcast = is_c, 0
        d.grid_re = d.grid_im = 0
        rho = torch.empty(B, dtype=torch.float64, device=self.device)
        grad = torch.empty((B, self.n_out, 2), dtype=torch.float64, device=self.device)
        lam_r = torch.view_as_real(lam)
        _lib.check(self._L.sdcgym_spectral_radius_grad(ctypes.byref(d), B, lam_r.data_ptr(), qd.data_ptr(),
                                                       rho.data_ptr(), grad.data_ptr(), self._stream()),
                   "sdcgym_spectral_radius_grad")
        self._keep = (lam_r, qd)
        return rho, torch.view_as_complex(grad)

    def value_and_grad(self, lams, outputs, convention="jax"):
        """(mean rho, d mean_rho / d outputs) - what ``jax.value_and_grad(loss)`` gives the reference's trainer for
        the loss itself (``dp_playground.py:1038-1073``).  For complex ``outputs`` the gradient follows
        ``convention``: 'jax' (``d/dx - i d/dy``) or 'torch' (its conjugate); for real outputs it is real."""
        torch = _torch()
        rho, g = self.radii_and_grads(lams, outputs)
        g = g / rho.numel()
        is_c = (outputs.is_complex() if isinstance(outputs, torch.Tensor) else np.iscomplexobj(outputs))
        if not is_c:
            g = g.real
        elif convention == "torch":
            g = g.conj()
        elif convention != "jax":
            raise ValueError("convention must be 'jax' or 'torch'")
        return self.mean(rho), g

    def differentiable(self, lams, outputs):
        """Mean spectral radius as a torch scalar that supports ``.backward()`` into ``outputs`` (a CUDA tensor with
        ``requires_grad``) - forward and backward both run the hand-written kernels."""
        return _SpectralRadiusFn.apply(outputs, self, lams)

    @_lib.on_device
    def _sum(self, rho):
        """deterministic fp64 sum (fixed reduction tree) as a CUDA scalar"""
        torch = _torch()
        out = torch.zeros(1, dtype=torch.float64, device=self.device)
        if rho.numel():
            scratch = torch.empty(self._L.sdcgym_sum_scratch_doubles(), dtype=torch.float64, device=self.device)
            _lib.check(self._L.sdcgym_sum_f64(rho.numel(), rho.data_ptr(), scratch.data_ptr(), out.data_ptr(),
                                              self._stream()), "sdcgym_sum_f64")
        return out[0]

    def mean(self, rho):
        return self._sum(rho) / rho.numel()

    @_lib.on_device
    def grid(self, n_re, n_im, lambda_real_interval, lambda_imag_interval, output=None, rows=None):
        """rho on the (n_re x n_im) tensor grid over the lambda box for ONE Q_delta parameter row (or the fixed
        ``prec``): lambdas are generated from the grid index on the device (0 bytes in, 8 bytes out per matrix).
        ``rows=(lo, hi)`` evaluates only the grid rows lo..hi-1 (real-axis index) - the shard of one rank.
        Returns a CUDA tensor (n_re, n_im) [(hi - lo, n_im) with ``rows``]."""
        torch = _torch()
        qd, is_c, _ = self._outputs_tensor(output, 1) if self.prec is None else (None, 0, 0)
        lo, hi = (0, int(n_re)) if rows is None else (int(rows[0]), int(rows[1]))
        if not 0 <= lo <= hi <= int(n_re):
            raise ValueError(f"rows={rows} outside the grid of {n_re} rows")
        d = self._desc
        d.qd_is_complex, d.qd_broadcast = is_c, 1
        d.grid_re, d.grid_im, d.grid_first = int(n_re), int(n_im), lo * int(n_im)
        d.re_lo, d.re_hi = float(lambda_real_interval[0]), float(lambda_real_interval[1])
        d.im_lo, d.im_hi = float(lambda_imag_interval[0]), float(lambda_imag_interval[1])
        N = (hi - lo) * int(n_im)
        rho = torch.empty(N, dtype=torch.float64, device=self.device)
        _lib.check(self._L.sdcgym_spectral_radius(ctypes.byref(d), N, None, None if qd is None else qd.data_ptr(),
                                                  rho.data_ptr(), self._stream()), "sdcgym_spectral_radius")
        d.grid_first = 0
        self._keep = (qd,)
        return rho.reshape(hi - lo, n_im)

    def grid_mean(self, n_re, n_im, lambda_real_interval, lambda_imag_interval, output=None):
        """Mean of rho over the grid with the rows sharded over the ranks of the default process group:
        every rank evaluates ``dist.shard_range(n_re)`` rows and ONE scalar is all-reduced.  Without an initialised
        process group this is the single-GPU mean.  Returns a CUDA scalar (the same value on every rank)."""
        from . import dist as _dist

        rank, world = _dist.world()
        lo, count = _dist.shard_range(int(n_re), rank, world)
        rho = self.grid(n_re, n_im, lambda_real_interval, lambda_imag_interval, output, rows=(lo, lo + count))
        total = _dist.all_reduce_sum(self._sum(rho.reshape(-1)).reshape(1))[0]
        return total / (int(n_re) * int(n_im))


def _make_autograd_fn():
    torch = _torch()

    class SpectralRadiusFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, outputs, loss, lams):
            value, g = loss.value_and_grad(lams, outputs, convention="torch")
            ctx.save_for_backward(g)
            ctx.out_shape, ctx.out_dtype = outputs.shape, outputs.dtype
            return value.reshape(())

        @staticmethod
        def backward(ctx, upstream):
            (g,) = ctx.saved_tensors
            return (g * upstream).reshape(ctx.out_shape).to(ctx.out_dtype), None, None

    return SpectralRadiusFn


class _LazyFn:
    _fn = None

    def apply(self, *a):
        if _LazyFn._fn is None:
            _LazyFn._fn = _make_autograd_fn()
        return _LazyFn._fn.apply(*a)


_SpectralRadiusFn = _LazyFn()

NormLoss = SpectralRadiusLoss  # the reference keeps this alias (dp_playground.py:233)


class ResidualLoss:
    """``dp_playground.py:235-258``: one sweep with the learned Q_delta; loss = mean ||r'||_inf.

    ``__call__(lams, outputs, Cs, u0s, us, old_residuals) -> (mean_norm, us', residuals')`` like the reference.
    ``Cs`` may be ``None`` (the system matrices are then formed from ``lams`` on the device instead of being read).
    """

    def __init__(self, M, dt, prec_type="diag", *, prec=None, Q=None, device=None):
        self._sr = SpectralRadiusLoss(M, dt, prec_type, prec=prec, Q=Q, device=device)
        self.M, self.dt, self.prec_type = self._sr.M, self._sr.dt, self._sr.prec_type
        self.Q, self.device = self._sr.Q, self._sr.device
        self._guard = self._sr._guard

    def _c128(self, x, shape):
        torch = _torch()
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.complex128))
        t = t.to(self.device).to(torch.complex128).reshape(shape).contiguous()
        return torch.view_as_real(t)

    @_lib.on_device
    def take_step(self, lams, outputs, Cs, u0s, us, old_residuals, with_grad=False):
        torch = _torch()
        sr, M = self._sr, self.M
        lam = self._c128(lams, (-1,))
        B = lam.shape[0]
        out = outputs.detach() if isinstance(outputs, torch.Tensor) else outputs
        qd, is_c, bc = sr._outputs_tensor(out, B)
        if bc:
            qd = qd.expand(B, *qd.shape[1:]).contiguous()
        d = sr._desc
        d.qd_is_complex, d.qd_broadcast = is_c, 0
        u0, u, r = (self._c128(x, (B, M)) for x in (u0s, us, old_residuals))
        C = None if Cs is None else self._c128(Cs, (B, M, M))
        u_out = torch.empty((B, M, 2), dtype=torch.float64, device=self.device)
        r_out = torch.empty((B, M, 2), dtype=torch.float64, device=self.device)
        norms = torch.empty(B, dtype=torch.float64, device=self.device)
        grad = torch.empty((B, sr.n_out, 2), dtype=torch.float64, device=self.device) if with_grad else None
        _lib.check(sr._L.sdcgym_residual_step(
            ctypes.byref(d), B, lam.data_ptr(), None if qd is None else qd.data_ptr(),
            None if C is None else C.data_ptr(), u0.data_ptr(), u.data_ptr(), r.data_ptr(), u_out.data_ptr(),
            r_out.data_ptr(), norms.data_ptr(), None if grad is None else grad.data_ptr(), sr._stream()),
            "sdcgym_residual_step")
        self._keep = (lam, qd, C, u0, u, r)
        res = (torch.view_as_complex(u_out), torch.view_as_complex(r_out), norms)
        return res + (torch.view_as_complex(grad),) if with_grad else res

    def __call__(self, lams, outputs, Cs, u0s, us, old_residuals):
        us_new, residuals, norms = self.take_step(lams, outputs, Cs, u0s, us, old_residuals)
        return self._sr.mean(norms), us_new, residuals

    def value_and_grad(self, lams, outputs, Cs, u0s, us, old_residuals, convention="jax"):
        """((loss, us', residuals'), d loss / d outputs): ``jax.value_and_grad(loss, has_aux=True)`` for the residual
        loss (``dp_playground.py:1038-1073``); gradient conventions as in ``SpectralRadiusLoss.value_and_grad``."""
        torch = _torch()
        us_new, residuals, norms, g = self.take_step(lams, outputs, Cs, u0s, us, old_residuals, with_grad=True)
        g = g / norms.numel()
        is_c = (outputs.is_complex() if isinstance(outputs, torch.Tensor) else np.iscomplexobj(outputs))
        if not is_c:
            g = g.real
        elif convention == "torch":
            g = g.conj()
        elif convention != "jax":
            raise ValueError("convention must be 'jax' or 'torch'")
        return (self._sr.mean(norms), us_new, residuals), g
