// certify.cuh - the per-env CERTIFICATE of the substitution sweep mode (SDCGYM_SWEEP_CERTIFIED).
//
// The substitution mode iterates  u <- u + Pinv (u0 - C u),  r = u0 - u + z (Q u)  with real-Q FMAs (fast_kernels.cuh):
// the same recurrence as the reference (sdc_env.py:229-231) but with a different rounding sequence.  Its iteration
// counts / done / err flags are only accepted where they are PROVABLY the reference's.  The proof obligation per env:
//
//   e_k = u~_k - u^ref_k  obeys  e_{k+1} = K e_k + eta_k,   K = I - Pinv C   (Pinv: the same bits in both modes),
//   so ||e_k||_w <= sum_j ||K^(k-1-j)||_w ||eta_j||_w            (w-weighted max norm, ||x||_w = max_m |x_m| / w_m)
//
// This file computes, per env and in FP32 (the FP32 pipe is idle in these FP64-bound kernels), rigorous upper bounds
//   kappa_j >= ||K^j||_w  for j = 1..B  (explicit powers X_j = fl32(X_{j-1} K^), rounding and input errors carried by a
//   scalar recurrence delta_j >= ||K^j - X_j||_w),  a geometric envelope  ||K^n||_w <= G theta^n  for ALL n
//   (theta^B >= kappa_B, G = max_{j<B} kappa_j / theta^j; sub-multiplicativity does the rest),
// and the coefficients of the local-error model  ||eta_k||_w <= a0 + a1 U + a2 R  and of the decision margin
//   | ||r~_k||inf - ||r^ref_k||inf | <= Gamma ||e_k||_w + b0 + b1 U
// (U, R = running bounds of max |u_j|, max |r_m|; derivation and constants: DESIGN.md 4b).  The explicit powers are what
// makes the bound usable: K = z Pinv (Q - Q_delta) is strongly non-normal for good preconditioners (rho(K) = 0.55 but
// rho(|K|) = 2.6 for the MIN diagonal), so any bound through |K| or ||K|| alone explodes within a dozen sweeps.
// The weights w are a few power-iteration steps towards the Perron vector of |K^2| (any positive w gives a valid norm,
// so they need no error analysis).
#pragma once
#include "exact_math.cuh"

namespace sdcgym {

constexpr int kCertPlanes = 8;  // theta, G, Gamma, a0, a1, a2, b0, b1  (float planes [8][ld])
constexpr int kCertPowers = 6;  // B: explicit powers K^1..K^B

struct Cert {
    float theta, G, gam, a0, a1, a2, b0, b1;
};

SDCGYM_HD float cert_inf() { return (float)INFINITY; }

// fp32 modulus, rounded up generously: s * rsqrt(s) with the 2-ulp hardware reciprocal square root on the device
// (relative error < 5e-7 in total), correctly rounded sqrtf on the host build
SDCGYM_HD float cabs_up(float re, float im) {
    // the 1e-36 keeps s > 0 (no guard needed; it only raises the bound) and is far below anything that matters
    const float s = fmaf(re, re, fmaf(im, im, 1e-36f));
#ifdef __CUDA_ARCH__
    return s * rsqrtf(s) * 1.000002f;  // (s = +Inf gives NaN: the certificate is then unusable, as it must be)
#else
    return sqrtf(s) * 1.000002f;
#endif
}

// Constants of the local error model, in units of eps = 2^-53:
//   c_ref  = 2 + sqrt(2) (M + 1)   reference residual  r = u0 - C^ @ u: rounded C entries (2), zgemv_t micro-kernel whose
//                                  longest rounding path has <= M + 1 roundings for M = 2..9 (exact_math.cuh
//                                  zgemv_rowdot: head chains of m1/2 + 2, tail of (M & 3) + 2, one joining add)
//   c_fast = M + 3                 substitution residual r = u0 - u + z (Q u)  (FMA chains of length M, 5 roundings)
template <int M>
struct CertModel {
    static constexpr float c_res = (2.0f + 1.41425f * (M + 1)) + (M + 3.0f);  // c_ref + c_fast
    static constexpr float c_upd_r = 8.5f;   // |Pinv| |r| terms: reference product + add (3.3), FMA update (2.5), eps |r| of both residuals (2.5)
    static constexpr float c_upd_u = 3.5f;   // |u| terms of the two u += delta roundings
};

// ---- generic part: powers, weights, envelope.  K (complex fp32, row-major M x M) and dK >= entrywise |K - K_true|
//      expressed as a bound on ||K - K_true||_w for ANY weights with max w = 1, min w >= wfloor. ----
template <int M, int B>
SDCGYM_HD void cert_envelope(const float (&Kr)[M * M], const float (&Ki)[M * M], const float (&dKrow)[M] /* sum_j |dK_mj| */,
                             float (&w)[M], float& theta, float& G) {
    constexpr float u32 = 5.9604645e-8f;        // 2^-24
    constexpr float g32 = (4.0f * M + 8.0f) * u32;  // complex dot of length M in fp32: <= 2M+2 roundings per component, x sqrt(2), generous
    constexpr float wfloor = 0.02f;
    // powers X_j = X_{j-1} K^ (in place, row by row)
    float Xr[M * M], Xi[M * M];
#pragma unroll
    for (int k = 0; k < M * M; k++) {
        Xr[k] = Kr[k];
        Xi[k] = Ki[k];
    }
    float kap[B + 1], iw[M];
    kap[0] = 1.0f;
    auto wnorm = [&](const float (&A)[M * M]) {  // max_m sum_j A_mj w_j / w_m  for a nonnegative matrix
        float best = 0.0f;
#pragma unroll
        for (int m = 0; m < M; m++) {
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < M; j++) s = fmaf(A[m * M + j], w[j], s);
            s = s * iw[m];
            best = s > best ? s : best;
        }
        return best;
    };
    auto square_step = [&]() {  // X <- X K, row by row
#pragma unroll
        for (int m = 0; m < M; m++) {
            float tr[M], ti[M];
#pragma unroll
            for (int j = 0; j < M; j++) {
                float sr = 0.0f, si = 0.0f;
#pragma unroll
                for (int l = 0; l < M; l++) {
                    sr = fmaf(Xr[m * M + l], Kr[l * M + j], fmaf(-Xi[m * M + l], Ki[l * M + j], sr));
                    si = fmaf(Xr[m * M + l], Ki[l * M + j], fmaf(Xi[m * M + l], Kr[l * M + j], si));
                }
                tr[j] = sr;
                ti[j] = si;
            }
#pragma unroll
            for (int j = 0; j < M; j++) {
                Xr[m * M + j] = tr[j];
                Xi[m * M + j] = ti[j];
            }
        }
    };
    // weights: power iteration on |K^2| (phase cancellation of two sweeps included), regularised, max-normalised
    square_step();
    float absX[M * M];
#pragma unroll
    for (int k = 0; k < M * M; k++) absX[k] = cabs_up(Xr[k], Xi[k]);
#pragma unroll
    for (int m = 0; m < M; m++) w[m] = 1.0f;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        float v[M], vmax = 0.0f;
#pragma unroll
        for (int m = 0; m < M; m++) {
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < M; j++) s = fmaf(absX[m * M + j], w[j], s);
            v[m] = s;
            vmax = s > vmax ? s : vmax;
        }
        if (!(vmax > 1e-30f) || !(vmax < 1e30f)) {
#pragma unroll
            for (int m = 0; m < M; m++) v[m] = 1.0f;
            vmax = 1.0f;
        }
        // additive regularisation (0.05 of the previous max weight = 1): keeps every weight away from zero
        const float inv = 1.0f / (vmax + 0.05f);
#pragma unroll
        for (int m = 0; m < M; m++) {
            float x = (v[m] + 0.05f) * inv;
            x = x > wfloor ? x : wfloor;
            w[m] = x < 1.0f ? x : 1.0f;
        }
    }
    // delta_1 = ||K^ - K||_w  <= max_m dKrow_m * (max w) / w_m
    float d1 = 0.0f;
#pragma unroll
    for (int m = 0; m < M; m++) {
        iw[m] = (1.0f / w[m]) * 1.000001f;
        const float t = dKrow[m] * iw[m];
        d1 = t > d1 ? t : d1;
    }
    d1 *= 1.0001f;
    // kappa_j >= ||K^j||_w.  With L_1 = K^ - K and L_i = X_{i-1} (K^ - K) + (fl32 rounding of the product X_{i-1} K^):
    //     X_j = K^j + sum_{i=1..j} L_i K^(j-i)   =>   ||K^j||_w <= ||X_j||_w + sum_i l_i kappa_(j-i),
    // l_1 = delta_1, l_i = ||X_{i-1}||_w (delta_1 + g32 ||K^||_w): the local errors are propagated by the (already
    // certified) norms of the LOWER powers, not by ||K||^(j-i) - which is what keeps fp32 good enough when
    // ||K|| = 7 but ||K^6|| = 0.07.
    float k1;
    {
        float absK[M * M];
#pragma unroll
        for (int k = 0; k < M * M; k++) absK[k] = cabs_up(Kr[k], Ki[k]);
        k1 = wnorm(absK) * 1.0001f;
    }
    float khat[B + 1], lerr[B + 1];
    khat[0] = 1.0f;
    khat[1] = k1;
    lerr[0] = 0.0f;
    lerr[1] = d1;
    kap[1] = k1 + d1;
    const float lstep = (d1 + g32 * k1) * 1.0001f;
    // rolled: one product body in the instruction stream (fully unrolled the kernel is 110 KB of straight-line code
    // that every warp runs once - it then waits on instruction fetch); the three short arrays go to local memory
#pragma unroll 1
    for (int j = 2; j <= B; j++) {
        if (j > 2) {
            square_step();
#pragma unroll
            for (int k = 0; k < M * M; k++) absX[k] = cabs_up(Xr[k], Xi[k]);
        }
        khat[j] = wnorm(absX) * 1.0001f;
        lerr[j] = khat[j - 1] * lstep;
        float acc = khat[j];
        for (int i = 1; i <= j; i++) acc = fmaf(lerr[i], kap[j - i], acc);
        kap[j] = acc * 1.0001f;
    }
    // geometric envelope  ||K^n||_w <= G theta^n
    float kb = kap[B];
    float th = (kb > 0.0f) ? exp2f(log2f(kb) * (1.0f / B)) * 1.00002f : 0.0f;
    th = th > 0.05f ? th : 0.05f;
    if (!(kb < 1e30f)) th = cert_inf();
    float g = 1.0f, tp = 1.0f;
#pragma unroll
    for (int j = 1; j < B; j++) {
        tp *= th;
        const float q = kap[j] / tp;
        g = q > g ? q : g;
    }
    theta = th;
    G = g * 1.0001f;
    if (!(G < 1e30f)) G = cert_inf();
}

// ---- diagonal Q_delta: Pinv = diag(pi_m) (the crecip results the sweeps use), K_mj = pi_m z q_mj + delta_mj (1 - pi_m) ----
template <int M, int B>
SDCGYM_HD Cert cert_diag(const double (&Q)[M * M], double zr_d, double zi_d, const double (&pir_d)[M], const double (&pii_d)[M]) {
    constexpr float u32 = 5.9604645e-8f;
    constexpr float eps = 1.1102230246251565e-16f;  // 2^-53
    const float zr = (float)zr_d, zi = (float)zi_d;
    const float zabs = cabs_up(zr, zi) * 1.000001f;
    float Kr[M * M], Ki[M * M], dKrow[M], pabs[M], Arow[M];
#pragma unroll
    for (int m = 0; m < M; m++) {
        const float pr = (float)pir_d[m], pi = (float)pii_d[m];
        pabs[m] = cabs_up(pr, pi) * 1.000001f;
        const float tr = pr * zr - pi * zi, ti = pr * zi + pi * zr;  // t_m = pi_m z
        const float tabs = pabs[m] * zabs;
        float arow = 0.0f;
#pragma unroll
        for (int j = 0; j < M; j++) {
            const float q = (float)Q[m * M + j];
            arow += fabsf(q);
            float kr = tr * q, ki = ti * q;
            if (j == m) {
                kr += 1.0f - pr;
                ki += -pi;
            }
            Kr[m * M + j] = kr;
            Ki[m * M + j] = ki;
        }
        Arow[m] = arow * 1.00001f;
        // |K^_mj - K_mj| <= 8 u32 |t_m| |q_mj| + delta_mj u32 (3 |pi_m| + 2)   (input roundings + fp32 evaluation);
        // doubled for the fp32 evaluation of the bound itself
        dKrow[m] = 16.0f * u32 * tabs * Arow[m] + 2.0f * u32 * (3.0f * pabs[m] + 2.0f);
    }
    float w[M];
    Cert c;
    cert_envelope<M, B>(Kr, Ki, dKrow, w, c.theta, c.G);
    // local error model and decision margin (units: absolute, eps folded in)
    float a1 = 0.0f, a0 = 0.0f, iw = 0.0f, gam = 0.0f, amax = 0.0f;
#pragma unroll
    for (int m = 0; m < M; m++) {
        const float pw = pabs[m] / w[m];
        const float zA1 = zabs * Arow[m] + 1.0f;
        const float t1 = pw * CertModel<M>::c_res * zA1;
        a1 = t1 > a1 ? t1 : a1;
        a0 = pw > a0 ? pw : a0;
        const float i1 = 1.0f / w[m];
        iw = i1 > iw ? i1 : iw;
        amax = zA1 > amax ? zA1 : amax;
        float s = w[m];
#pragma unroll
        for (int j = 0; j < M; j++) s += zabs * fabsf((float)Q[m * M + j]) * 1.00001f * w[j];
        gam = s > gam ? s : gam;
    }
    constexpr float infl = 1.001f;
    c.a1 = eps * (a1 + CertModel<M>::c_upd_u * iw) * infl;
    c.a0 = eps * a0 * CertModel<M>::c_res * infl;
    c.a2 = eps * a0 * CertModel<M>::c_upd_r * infl;
    c.gam = gam * infl;
    c.b1 = eps * CertModel<M>::c_res * amax * infl;
    c.b0 = eps * CertModel<M>::c_res * infl;
    // anything not finite makes every decision of this env ambiguous (the exact kernel takes it)
    if (!(c.a1 < 1e30f) || !(c.gam < 1e30f) || !(c.a0 < 1e30f) || !(c.b1 < 1e30f)) {
        c.theta = cert_inf();
        c.G = cert_inf();
    }
    return c;
}

}  // namespace sdcgym
