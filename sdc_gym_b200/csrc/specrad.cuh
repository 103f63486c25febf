// specrad.cuh - batched spectral radius of the SDC iteration matrix, one matrix per thread.
//
//   rho( lam*dt * inv(I - lam*dt*Qd) @ (Q - Qd) )        dp_playground.py:216-228, sdc_env.py:421-425
//
// The reference evaluates this with LAPACK (inv + matmul + zgeev) per sample.  Here each thread
//   1. forms K = z * P^{-1} (Q - Qd) by forward substitution (all supported Q_delta are lower triangular:
//      diag / lower_diag / lower_tri / strictly_lower_tri layouts and the fixed LU^T, MIN, EE, zeros),
//   2. reduces K to upper Hessenberg form with Householder reflectors,
//   3. runs the explicitly shifted complex QR iteration (Wilkinson shift, Givens rotations, deflation from
//      the bottom) and keeps max |eigenvalue|.
// Everything is unrolled over compile-time indices with run-time predicates so the M x M matrix stays in
// registers (M <= 5) instead of local memory.  Power iteration is unusable here: |mu_2/mu_1| has median
// 0.95 on the lambda box (hard part 4).  Target accuracy: 1e-10 relative against LAPACK.
//
// Unlike the env kernels nothing here has to follow a particular rounding sequence, so this file is compiled
// with FMA contraction enabled.
#pragma once
#include "exact_math.cuh"

namespace sdcgym {

struct C2 {
    double r, i;
};
SDCGYM_HD C2 c_add(C2 a, C2 b) { return C2{a.r + b.r, a.i + b.i}; }
SDCGYM_HD C2 c_sub(C2 a, C2 b) { return C2{a.r - b.r, a.i - b.i}; }
SDCGYM_HD C2 c_mul(C2 a, C2 b) { return C2{a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
SDCGYM_HD C2 c_conj(C2 a) { return C2{a.r, -a.i}; }
SDCGYM_HD C2 c_scale(C2 a, double s) { return C2{a.r * s, a.i * s}; }
SDCGYM_HD double c_abs2(C2 a) { return a.r * a.r + a.i * a.i; }
SDCGYM_HD double c_abs1(C2 a) { return fabs(a.r) + fabs(a.i); }
SDCGYM_HD C2 c_div(C2 a, C2 b) {  // Smith's algorithm
    C2 q;
    if (fabs(b.r) >= fabs(b.i)) {
        double t = b.i / b.r, d = b.r + b.i * t;
        q.r = (a.r + a.i * t) / d;
        q.i = (a.i - a.r * t) / d;
    } else {
        double t = b.r / b.i, d = b.r * t + b.i;
        q.r = (a.r * t + a.i) / d;
        q.i = (a.i * t - a.r) / d;
    }
    return q;
}
SDCGYM_HD C2 c_sqrt(C2 a) {  // principal square root
    double mag = sqrt(c_abs2(a));
    if (mag == 0.0) return C2{0.0, 0.0};
    if (a.r >= 0.0) {
        double re = sqrt(0.5 * (mag + a.r));
        return C2{re, a.i / (2.0 * re)};
    }
    double im = sqrt(0.5 * (mag - a.r));
    return C2{fabs(a.i) / (2.0 * im), a.i >= 0.0 ? im : -im};
}

// eigenvalue modulus maximum of the M x M complex matrix H (row-major, destroyed)
template <int M>
SDCGYM_HD double max_abs_eig(C2 (&H)[M * M]) {
#define H_(i, j) H[(i) * M + (j)]
    // ---- Householder reduction to upper Hessenberg form ----
#pragma unroll
    for (int k = 0; k < M - 2; k++) {
        double nrm2 = 0.0;
#pragma unroll
        for (int i = k + 1; i < M; i++) nrm2 += c_abs2(H_(i, k));
        double tail2 = nrm2 - c_abs2(H_(k + 1, k));
        if (tail2 > 0.0 && nrm2 > 0.0) {
            double nrm = sqrt(nrm2);
            C2 x0 = H_(k + 1, k);
            double ax0 = sqrt(c_abs2(x0));
            C2 phase = (ax0 == 0.0) ? C2{1.0, 0.0} : C2{x0.r / ax0, x0.i / ax0};
            C2 alpha = c_scale(phase, -nrm);  // x -> alpha e1
            C2 v[M];
#pragma unroll
            for (int i = 0; i < M; i++) v[i] = (i > k) ? H_(i, k) : C2{0.0, 0.0};
            v[k + 1] = c_sub(v[k + 1], alpha);
            double vn2 = 0.0;
#pragma unroll
            for (int i = k + 1; i < M; i++) vn2 += c_abs2(v[i]);
            double beta = 2.0 / vn2;
            // H <- (I - beta v v^H) H
#pragma unroll
            for (int j = 0; j < M; j++) {
                C2 s{0.0, 0.0};
#pragma unroll
                for (int i = k + 1; i < M; i++) s = c_add(s, c_mul(c_conj(v[i]), H_(i, j)));
                s = c_scale(s, beta);
#pragma unroll
                for (int i = k + 1; i < M; i++) H_(i, j) = c_sub(H_(i, j), c_mul(v[i], s));
            }
            // H <- H (I - beta v v^H)
#pragma unroll
            for (int i = 0; i < M; i++) {
                C2 s{0.0, 0.0};
#pragma unroll
                for (int j = k + 1; j < M; j++) s = c_add(s, c_mul(H_(i, j), v[j]));
                s = c_scale(s, beta);
#pragma unroll
                for (int j = k + 1; j < M; j++) H_(i, j) = c_sub(H_(i, j), c_mul(s, c_conj(v[j])));
            }
#pragma unroll
            for (int i = k + 2; i < M; i++) H_(i, k) = C2{0.0, 0.0};
        }
    }

    // ---- shifted QR on the Hessenberg matrix, deflating from the bottom ----
    const double eps = 2.220446049250313e-16;
    double rho = 0.0;
    int hi = M - 1;
    int its = 0;
    bool failed = false;
    for (int guard = 0; guard < 40 * M && hi > 0; guard++) {
        // negligible sub-diagonal at the bottom of the active block?
        C2 sub{0.0, 0.0}, d0{0.0, 0.0}, d1{0.0, 0.0}, b01{0.0, 0.0};
#pragma unroll
        for (int k = 1; k < M; k++)
            if (k == hi) {
                sub = H_(k, k - 1);
                d0 = H_(k - 1, k - 1);
                d1 = H_(k, k);
                b01 = H_(k - 1, k);
            }
        double scale = c_abs1(d0) + c_abs1(d1);
        if (c_abs1(sub) <= eps * scale || c_abs1(sub) == 0.0) {
            double a = sqrt(c_abs2(d1));
            rho = a > rho ? a : rho;
            hi--;
            its = 0;
            continue;
        }
        if (its >= 60) {
            failed = true;
            break;
        }
        // Wilkinson shift: eigenvalue of [[d0, b01], [sub, d1]] closer to d1
        C2 mu;
        if (its == 10 || its == 20) {
            mu = C2{d1.r + fabs(sub.r) + fabs(sub.i), d1.i};  // exceptional shift
        } else {
            C2 half = c_scale(c_sub(d0, d1), 0.5);
            C2 bc = c_mul(b01, sub);
            C2 disc = c_sqrt(c_add(c_mul(half, half), bc));
            C2 p = c_add(half, disc), q = c_sub(half, disc);
            C2 den = (c_abs2(p) >= c_abs2(q)) ? p : q;
            mu = (c_abs2(den) == 0.0) ? d1 : c_sub(d1, c_div(bc, den));
        }
        its++;
#pragma unroll
        for (int k = 0; k < M; k++)
            if (k <= hi) H_(k, k) = c_sub(H_(k, k), mu);
        // QR factorisation by Givens rotations (rows), rotations kept for the RQ product
        C2 gc[M], gs[M];
#pragma unroll
        for (int k = 0; k < M - 1; k++) {
            gc[k] = C2{1.0, 0.0};
            gs[k] = C2{0.0, 0.0};
            if (k < hi) {
                C2 x = H_(k, k), y = H_(k + 1, k);
                double r = sqrt(c_abs2(x) + c_abs2(y));
                if (r > 0.0) {
                    double ir = 1.0 / r;
                    gc[k] = c_scale(x, ir);
                    gs[k] = c_scale(y, ir);
                }
                C2 cc = c_conj(gc[k]), cs = c_conj(gs[k]);
#pragma unroll
                for (int j = k; j < M; j++) {
                    if (j <= hi) {
                        C2 a = H_(k, j), b = H_(k + 1, j);
                        H_(k, j) = c_add(c_mul(cc, a), c_mul(cs, b));
                        H_(k + 1, j) = c_sub(c_mul(gc[k], b), c_mul(gs[k], a));
                    }
                }
            }
        }
        // RQ: apply the conjugate-transposed rotations to the columns
#pragma unroll
        for (int k = 0; k < M - 1; k++) {
            if (k < hi) {
                C2 cc = c_conj(gc[k]), cs = c_conj(gs[k]);
#pragma unroll
                for (int i = 0; i <= k + 1; i++) {
                    C2 a = H_(i, k), b = H_(i, k + 1);
                    H_(i, k) = c_add(c_mul(a, gc[k]), c_mul(b, gs[k]));
                    H_(i, k + 1) = c_sub(c_mul(b, cc), c_mul(a, cs));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < M; k++)
            if (k <= hi) H_(k, k) = c_add(H_(k, k), mu);
    }
    if (failed || hi > 0) return d_nan();
    double a = sqrt(c_abs2(H_(0, 0)));
    return a > rho ? a : rho;
#undef H_
}

// K = z * P^{-1} (Q - Qd) with P = I - z Qd lower triangular.  Qd: row-major dense complex (upper part ignored).
template <int M>
SDCGYM_HD double spectral_radius_one(const double* Q, double zr, double zi, const C2 (&Qd)[M * M]) {
    const C2 z{zr, zi};
    C2 X[M * M];
#pragma unroll
    for (int i = 0; i < M; i++) {
        C2 pii = c_sub(C2{1.0, 0.0}, c_mul(z, Qd[i * M + i]));
        C2 inv = c_div(C2{1.0, 0.0}, pii);
#pragma unroll
        for (int c = 0; c < M; c++) {
            C2 acc{0.0, 0.0};
#pragma unroll
            for (int j = 0; j < i; j++) acc = c_add(acc, c_mul(Qd[i * M + j], X[j * M + c]));
            C2 b = C2{Q[i * M + c], 0.0};
            if (c <= i) b = c_sub(b, Qd[i * M + c]);
            X[i * M + c] = c_mul(c_add(b, c_mul(z, acc)), inv);
        }
    }
#pragma unroll
    for (int k = 0; k < M * M; k++) X[k] = c_mul(z, X[k]);
    return max_abs_eig<M>(X);
}

}  // namespace sdcgym
