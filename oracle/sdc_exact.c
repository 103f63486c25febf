/*
 * oracle/sdc_exact.c - TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the arithmetic the
 * reference executes for one SDC env step (sdc_gym/envs/sdc_env.py), rounding for rounding.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * file's shared object (oracle/_build/libsdc_oracle.so).  The product path (sdc_gym_b200/) never does.
 *
 * The reference is numpy complex128 code; its bits are decided by three third-party pieces that are not
 * in /root/reference: numpy's complex ufunc loops (multiply, absolute), OpenBLAS 0.3.30's zgemv_t kernel
 * behind numpy's `@`, and LAPACK zgesv (OpenBLAS zgetf2 + zgetrs/ztrsm) behind np.linalg.inv.  Their
 * rounding sequences are restated here.  Parity pin: tests/test_oracle_*.py
 * check this file (a) against golden vectors produced by running the unmodified reference env
 * (tests/golden/make_golden.py through oracle/ref_loader.py) and (b) live against numpy on the host
 * ("BLAS fingerprint").  `variant` selects the OpenBLAS run-time core: 0 = SkylakeX (scalar C tails
 * compiled with FMA contraction), 1 = Haswell (no contraction in the scalar tails).
 *
 * Layouts are numpy-natural: complex = interleaved (re, im) doubles, env-major.
 *   lam[N][2], u[N][M][2], r[N][M][2], action[N][A] (real) or [N][A][2] (complex), Q[M][M] real row-major.
 *
 * Reference lines restated (sdc_env.py):
 *   C = eye(M) - lam*dt*Q ............................ :302-304   -> build_C
 *   d = interp(a, (-1,1), (0,1)) ...................... :125-132   -> scale_action
 *   Qd from action / fixed matrix ..................... :134-191   -> build_P (layouts: dp_playground.py:194-207)
 *   Pinv = inv(eye(M) - lam*dt*Qd) .................... :193-201   -> cinv (zgetf2 + zgetrs emulation)
 *   u += Pinv @ residual .............................. :229, :516 -> zgemv_rowdot
 *   residual = u0 - C @ u ............................. :203-204
 *   norm = max_m |v_m| ................................ :206-207   -> np_cabs / inf_norm
 *   sdc-v0 loop / err / done .......................... :209-273   -> sdc_oracle_step_v0
 *   sdc-v1 step ....................................... :507-572   -> sdc_oracle_step_v1
 *   rewards ........................................... :334-463   -> reward_func
 *   reset ............................................. :306-332   -> sdc_oracle_reset
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -mfma (see oracle/Makefile).  -ffp-contract=off makes every
 * `a*b+c` below two roundings; every single-rounding fused op is written as fma().
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXM 16

typedef struct { double re, im; } cplx;

enum { PREC_DIAG = 0, PREC_LOWER_DIAG = 1, PREC_LOWER_TRI = 2, PREC_STRICTLY_LOWER_TRI = 3, PREC_FIXED = 4 };
enum { REW_ITERATION_ONLY = 0, REW_RESIDUAL_CHANGE = 1, REW_GAUSS_KERNEL = 2, REW_FAST_CONVERGENCE = 3,
       REW_SMOOTH_FAST_CONVERGENCE = 4, REW_SMOOTHER_FAST_CONVERGENCE = 5, REW_SPECTRAL_RADIUS = 6 };

/* ---- numpy complex multiply loop: a*b = (fms(ar,br, ai*bi), fma(ar,bi, ai*br))  (Appendix A step 0) ---- */
static inline cplx cmul_np(cplx a, cplx b) {
    cplx c;
    c.re = fma(a.re, b.re, -(a.im * b.im));
    c.im = fma(a.re, b.im, a.im * b.re);
    return c;
}
static inline cplx cmul_unfused(cplx a, cplx b) {
    cplx c;
    c.re = a.re * b.re - a.im * b.im;
    c.im = a.re * b.im + a.im * b.re;
    return c;
}
/* product as compiled in OpenBLAS' scalar C kernels: contracted on SkylakeX, plain on Haswell */
static inline cplx cmul_blas(cplx a, cplx b, int variant) {
    return variant == 0 ? cmul_np(a, b) : cmul_unfused(a, b);
}
static inline cplx cadd(cplx a, cplx b) { cplx c = { a.re + b.re, a.im + b.im }; return c; }
static inline cplx csub(cplx a, cplx b) { cplx c = { a.re - b.re, a.im - b.im }; return c; }
static inline cplx cneg(cplx a) { cplx c = { -a.re, -a.im }; return c; }

/* ---- numpy |z| loop (AVX512F build): L*sqrt(fma(s/L, s/L, 1)), L = max(|re|,|im|)  (Appendix A step 9) ---- */
double sdc_oracle_cabs(double re, double im) {
    double a = fabs(re), b = fabs(im);
    if (isnan(a) || isnan(b)) {
        if (isinf(a) || isinf(b)) return INFINITY;
        return NAN;
    }
    if (isinf(a) || isinf(b)) return INFINITY;
    double L = a > b ? a : b, s = a > b ? b : a;
    if (L == 0.0) return 0.0;
    double t = s / L;
    return L * sqrt(fma(t, t, 1.0));
}
/* np.linalg.norm(v, inf) = abs(v).max(), NaN-propagating */
static double inf_norm(const cplx *v, int M) {
    double m = -INFINITY;
    int nan = 0;
    for (int i = 0; i < M; i++) {
        double a = sdc_oracle_cabs(v[i].re, v[i].im);
        if (isnan(a)) nan = 1;
        if (a > m) m = a;
    }
    return nan ? NAN : m;
}

/* ---- OpenBLAS complex reciprocal as used by ztrsm/zgetf2 (Appendix A step 6) ---- */
static inline cplx crecip(cplx p, int fused) {
    cplx inv;
    if (fabs(p.re) >= fabs(p.im)) {
        double t = p.im / p.re;
        double den = 1.0 / (p.re * (fused ? fma(t, t, 1.0) : (1.0 + t * t)));
        inv.re = den;
        inv.im = -t * den;
    } else {
        double t = p.re / p.im;
        double den = 1.0 / (p.im * (fused ? fma(t, t, 1.0) : (1.0 + t * t)));
        inv.re = t * den;
        inv.im = -den;
    }
    return inv;
}

/* ---- one row of numpy `A @ x` for C-contiguous complex (M,M)@(M,): OpenBLAS zgemv_t  (Appendix A step 3) ---- */
static cplx zgemv_rowdot(const cplx *a, const cplx *x, int M, int variant) {
    int m3 = M & 3, m1 = M - m3;
    cplx y = { 0.0, 0.0 };
    if (m1 > 0) {
        double E[4] = { 0, 0, 0, 0 }, O[4] = { 0, 0, 0, 0 };
        for (int j = 0; j < m1; j++) {
            double *S = (j & 1) ? O : E;
            S[0] = fma(x[j].re, a[j].re, S[0]);
            S[1] = fma(x[j].re, a[j].im, S[1]);
            S[2] = fma(x[j].im, a[j].re, S[2]);
            S[3] = fma(x[j].im, a[j].im, S[3]);
        }
        double lo_re = E[0] - E[3], lo_im = E[1] + E[2];
        double hi_re = O[0] - O[3], hi_im = O[1] + O[2];
        y.re = 0.0 + (hi_re + lo_re);
        y.im = 0.0 + (hi_im + lo_im);
    }
    if (m3) {
        cplx t = cmul_blas(a[m1], x[m1], variant);
        for (int j = m1 + 1; j < M; j++) {
            cplx e = cmul_blas(a[j], x[j], variant);
            t.re = e.re + t.re;
            t.im = e.im + t.im;
        }
        y.re = t.re + y.re;
        y.im = t.im + y.im;
    }
    return y;
}
void sdc_oracle_zgemv(int M, const double *A, const double *x, double *y, int variant) {
    for (int i = 0; i < M; i++) {
        cplx v = zgemv_rowdot((const cplx *)A + (size_t)i * M, (const cplx *)x, M, variant);
        y[2 * i] = v.re;
        y[2 * i + 1] = v.im;
    }
}

/* ================= np.linalg.inv for M <= 9: zgetf2 (left-looking, pivoted) + zgetrs  (Appendix A.2) ========= */

/* column-major helpers: A(i,j) = A[i + j*n] */
#define AT(A, i, j, n) ((A)[(i) + (size_t)(j) * (n)])

static void zgetf2_emul(cplx *A, int n, int *ipiv, int variant) {
    cplx b[MAXM];
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) b[i] = AT(A, i, j, n);
        /* 1. apply earlier row swaps to this column */
        for (int i = 0; i < j; i++) {
            int p = ipiv[i];
            if (p != i) { cplx t = b[i]; b[i] = b[p]; b[p] = t; }
        }
        /* 2. ztrsv_NLU on b[0:j] */
        for (int i = 0; i < j; i++) {
            cplx alpha = cneg(b[i]);
            for (int k = i + 1; k < j; k++) {
                cplx pr = cmul_blas(alpha, AT(A, k, i, n), variant);
                b[k].re = pr.re + b[k].re;
                b[k].im = pr.im + b[k].im;
            }
        }
        /* 3. zgemv_n: b[j:] -= A[j:, 0:j] @ b[0:j] */
        if (j > 0) {
            int rows = n - j, r4 = rows & ~3;
            for (int ii = 0; ii < r4; ii++) {
                int i = j + ii;
                cplx ybuf = { 0.0, 0.0 };
                int c = 0;
                int nblk4 = j >> 2;
                for (int blk = 0; blk < nblk4 + ((j & 2) ? 1 : 0) + ((j & 1) ? 1 : 0); blk++) {
                    int w = blk < nblk4 ? 4 : ((blk == nblk4 && (j & 2)) ? 2 : 1);
                    double S1 = 0, S2 = 0, S3 = 0, S4 = 0;
                    for (int q = 0; q < w; q++, c++) {
                        cplx a = AT(A, i, c, n), x = b[c];
                        if (q == 0) {
                            S1 = x.re * a.re; S2 = x.re * a.im; S3 = x.im * a.re; S4 = x.im * a.im;
                        } else {
                            S1 = fma(x.re, a.re, S1); S2 = fma(x.re, a.im, S2);
                            S3 = fma(x.im, a.re, S3); S4 = fma(x.im, a.im, S4);
                        }
                    }
                    ybuf.re = ybuf.re + (S1 - S4);
                    ybuf.im = ybuf.im + (S2 + S3);
                }
                b[i].re = b[i].re + (-ybuf.re);
                b[i].im = b[i].im + (-ybuf.im);
            }
            for (int ii = r4; ii < rows; ii++) {
                int i = j + ii;
                cplx t = { 0.0, 0.0 };
                for (int c = 0; c < j; c++) {
                    cplx pr = cmul_blas(AT(A, i, c, n), b[c], variant);
                    t.re = t.re + pr.re;
                    t.im = t.im + pr.im;
                }
                b[i].re = (-t.re) + b[i].re;
                b[i].im = (-t.im) + b[i].im;
            }
        }
        /* 4. pivot = first row >= j maximising |re| + |im| */
        int p = j;
        double best = fabs(b[j].re) + fabs(b[j].im);
        for (int i = j + 1; i < n; i++) {
            double v = fabs(b[i].re) + fabs(b[i].im);
            if (v > best) { best = v; p = i; }
        }
        ipiv[j] = p;
        for (int i = 0; i < n; i++) AT(A, i, j, n) = b[i];
        if (p != j) {
            for (int c = 0; c <= j; c++) { cplx t = AT(A, j, c, n); AT(A, j, c, n) = AT(A, p, c, n); AT(A, p, c, n) = t; }
        }
        /* 5. scale sub-column by the (always unfused) reciprocal of the pivot, unfused product (zscal) */
        cplx inv = crecip(AT(A, j, j, n), 0);
        for (int k = j + 1; k < n; k++) AT(A, k, j, n) = cmul_unfused(inv, AT(A, k, j, n));
    }
}

/* One ztrsm (left, no-trans): unit lower forward (upper=0) or non-unit upper backward (upper=1).
 * B (n x n column-major) is overwritten by the solution.  Tiles: rows 4,4,..,(2),(1); columns 2,2,..,(1). */
static void ztrsm_emul(const cplx *A, cplx *B, int n, int upper, int variant) {
    /* row tiles in storage order */
    int rt_start[MAXM], rt_size[MAXM], nrt = 0;
    {
        int i = 0;
        while (n - i >= 4) { rt_start[nrt] = i; rt_size[nrt++] = 4; i += 4; }
        if (n - i >= 2) { rt_start[nrt] = i; rt_size[nrt++] = 2; i += 2; }
        if (n - i >= 1) { rt_start[nrt] = i; rt_size[nrt++] = 1; i += 1; }
    }
    /* visiting order */
    int order[MAXM], no = 0;
    if (!upper) {
        for (int t = 0; t < nrt; t++) order[no++] = t;
    } else {
        /* remainder tiles first (1-tile, then 2-tile), then 4-tiles bottom to top */
        for (int t = nrt - 1; t >= 0; t--) if (rt_size[t] == 1) order[no++] = t;
        for (int t = nrt - 1; t >= 0; t--) if (rt_size[t] == 2) order[no++] = t;
        for (int t = nrt - 1; t >= 0; t--) if (rt_size[t] == 4) order[no++] = t;
    }
    cplx invd[MAXM];
    if (upper) for (int i = 0; i < n; i++) invd[i] = crecip(AT(A, i, i, n), variant == 0);

    for (int col0 = 0; col0 < n;) {
        int cw = (n - col0 >= 2) ? 2 : 1;
        for (int oi = 0; oi < no; oi++) {
            int t = order[oi], r0 = rt_start[t], rs = rt_size[t];
            for (int cc = 0; cc < cw; cc++) {
                int col = col0 + cc;
                /* (a) update with all already-solved rows */
                int p_lo = upper ? r0 + rs : 0, p_hi = upper ? n : r0;
                for (int ii = 0; ii < rs; ii++) {
                    int i = r0 + ii;
                    if (p_hi > p_lo) {
                        cplx val;
                        if (rs == 4) {
                            double Srr = 0, Sii = 0, Sri = 0, Sir = 0;
                            for (int p = p_lo; p < p_hi; p++) {
                                cplx a = AT(A, i, p, n), bb = AT(B, p, col, n);
                                Srr = fma(a.re, bb.re, Srr); Sii = fma(a.im, bb.im, Sii);
                                Sri = fma(a.re, bb.im, Sri); Sir = fma(a.im, bb.re, Sir);
                            }
                            val.re = Srr - Sii; val.im = Sir + Sri;
                        } else {
                            double re = 0, im = 0;
                            for (int p = p_lo; p < p_hi; p++) {
                                cplx a = AT(A, i, p, n), bb = AT(B, p, col, n);
                                re = fma(bb.re, a.re, -fma(bb.im, a.im, -re));   /* fms(br,ar, fms(bi,ai,re)) */
                                im = fma(bb.re, a.im, fma(bb.im, a.re, im));
                            }
                            val.re = re; val.im = im;
                        }
                        AT(B, i, col, n) = csub(AT(B, i, col, n), val);
                    }
                }
                /* (b) in-tile solve */
                for (int s = 0; s < rs; s++) {
                    int i = upper ? r0 + rs - 1 - s : r0 + s;
                    cplx ccv = upper ? cmul_blas(invd[i], AT(B, i, col, n), variant) : AT(B, i, col, n);
                    AT(B, i, col, n) = ccv;
                    for (int s2 = s + 1; s2 < rs; s2++) {
                        int k = upper ? r0 + rs - 1 - s2 : r0 + s2;
                        AT(B, k, col, n) = csub(AT(B, k, col, n), cmul_blas(ccv, AT(A, k, i, n), variant));
                    }
                }
            }
        }
        col0 += cw;
    }
}

/* Pinv = np.linalg.inv(P), P and Pinv row-major complex (M x M) */
void sdc_oracle_cinv(int M, const double *P_, double *Pinv_, int variant) {
    const cplx *P = (const cplx *)P_;
    cplx *Pinv = (cplx *)Pinv_;
    cplx A[MAXM * MAXM], B[MAXM * MAXM];
    int ipiv[MAXM];
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) AT(A, i, j, M) = P[i * M + j];
    zgetf2_emul(A, M, ipiv, variant);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) { AT(B, i, j, M).re = (i == j) ? 1.0 : 0.0; AT(B, i, j, M).im = 0.0; }
    for (int i = 0; i < M; i++) {
        int p = ipiv[i];
        if (p != i)
            for (int c = 0; c < M; c++) { cplx t = AT(B, i, c, M); AT(B, i, c, M) = AT(B, p, c, M); AT(B, p, c, M) = t; }
    }
    ztrsm_emul(A, B, M, 0, variant);
    ztrsm_emul(A, B, M, 1, variant);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) Pinv[i * M + j] = AT(B, i, j, M);
}

/* ========================================= env pieces ========================================= */

/* C = eye(M) - (lam*dt)*Q   (sdc_env.py:302-304; Appendix A steps 1-2) */
static void build_C(int M, const double *Q, cplx z, cplx *C) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            double q = Q[i * M + j];
            C[i * M + j].re = (i == j ? 1.0 : 0.0) - z.re * q;
            C[i * M + j].im = 0.0 - z.im * q;
        }
}

/* np.interp(a, (-1,1), (0,1)) (sdc_env.py:129).  numpy: x<=-1 -> 0, x>=1 -> 1, else slope*(x-(-1))+0 */
static inline double scale_action(double a) {
    if (isnan(a)) return NAN;
    if (a <= -1.0) return 0.0;
    if (a >= 1.0) return 1.0;
    return 0.5 * (a - (-1.0)) + 0.0;
}

int sdc_oracle_num_actions(int M, int prec_type) {
    switch (prec_type) {
    case PREC_DIAG: return M;
    case PREC_LOWER_DIAG: return M - 1;
    case PREC_LOWER_TRI: return M * (M + 1) / 2;
    case PREC_STRICTLY_LOWER_TRI: return M * (M - 1) / 2;
    default: return 0;
    }
}

/* Qd (dense complex) from the scaled action (sdc_env.py:134-140 for diag; dp_playground.py:194-207 layouts
 * for the others: row-major tril order) or from the fixed real matrix (LU / min / EE / zeros). */
static void build_Qd(int M, int prec_type, const cplx *d, const double *Qd_fixed, cplx *Qd) {
    for (int k = 0; k < M * M; k++) { Qd[k].re = 0.0; Qd[k].im = 0.0; }
    int k = 0;
    switch (prec_type) {
    case PREC_DIAG: for (int i = 0; i < M; i++) Qd[i * M + i] = d[i]; break;
    case PREC_LOWER_DIAG: for (int i = 1; i < M; i++) Qd[i * M + i - 1] = d[i - 1]; break;
    case PREC_LOWER_TRI: for (int i = 0; i < M; i++) for (int j = 0; j <= i; j++) Qd[i * M + j] = d[k++]; break;
    case PREC_STRICTLY_LOWER_TRI: for (int i = 1; i < M; i++) for (int j = 0; j < i; j++) Qd[i * M + j] = d[k++]; break;
    default: for (int q = 0; q < M * M; q++) Qd[q].re = Qd_fixed[q]; break;
    }
}

/* P = eye(M) - (lam*dt)*Qd  (sdc_env.py:198-200).  Real Qd: numpy multiplies the complex scalar with a
 * float array (cast to complex, imaginary part +0) - the ufunc loop is the same cmul. */
/* use_doubles=False: Qdmat is float32 / complex64 and lam*dt a Python complex, so numpy evaluates the product
 * in complex64 (operands rounded to float32, same fused loop form in float arithmetic); `eye(M) - ...`
 * promotes the result to complex128 exactly. */
static cplx cmul_np_f32(cplx a, cplx b) {
    float ar = (float)a.re, ai = (float)a.im, br = (float)b.re, bi = (float)b.im;
    float t1 = ai * bi, t2 = ai * br;
    cplx c;
    c.re = (double)fmaf(ar, br, -t1);
    c.im = (double)fmaf(ar, bi, t2);
    return c;
}

static void build_P(int M, cplx z, const cplx *Qd, cplx *P, int f32) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            cplx zq = f32 ? cmul_np_f32(z, Qd[i * M + j]) : cmul_np(z, Qd[i * M + j]);
            P[i * M + j].re = (i == j ? 1.0 : 0.0) - zq.re;
            P[i * M + j].im = 0.0 - zq.im;
        }
}

typedef struct {
    int strategy;          /* REW_* */
    double step_penalty, residual_weight, norm_factor, restol;
    int max_iters;
} reward_cfg;

/* norm of (v * norm_factor): numpy multiplies the complex array by the python scalar */
static double scaled_inf_norm(const cplx *v, int M, double nf) {
    cplx t[MAXM];
    cplx s = { nf, 0.0 };
    for (int i = 0; i < M; i++) t[i] = cmul_np(v[i], s);
    return inf_norm(t, M);
}

/* reward_func (sdc_env.py:427-463) except 'spectral_radius' (host side, needs eigvals).  Returns NAN and
 * sets *domain_err when math.log would raise ValueError (log(0)). */
static double reward_func(const reward_cfg *cfg, int M, const cplx *old_res, const cplx *res, const cplx *init_res,
                          int converged, int steps, int *domain_err) {
    double sp = cfg->step_penalty;
    switch (cfg->strategy) {
    case REW_ITERATION_ONLY:
        return -(double)steps * sp;
    case REW_RESIDUAL_CHANGE: {
        double a = scaled_inf_norm(old_res, M, cfg->norm_factor);
        double b = scaled_inf_norm(res, M, cfg->norm_factor);
        double c = scaled_inf_norm(init_res, M, cfg->norm_factor);
        double e = cfg->restol * cfg->norm_factor;
        if (!(a > 0) || !(b > 0) || !(c > 0) || !(e > 0)) { if (domain_err) *domain_err = 1; }
        double rew = fabs((log(a) - log(b)) / (log(c) - log(e)));
        rew *= cfg->residual_weight;
        rew -= (double)steps * sp;
        return rew;
    }
    case REW_GAUSS_KERNEL: {
        double nr = inf_norm(res, M);
        double ginv = 1.0 / cfg->restol;
        double x = nr * ginv;
        double gd = 0 + 1 * exp(-(x * x) / 2);
        double extra = 1;
        if (converged) { double k = (double)(cfg->max_iters + 1 - steps); extra = k * k * 10; }
        return gd * extra;
    }
    case REW_FAST_CONVERGENCE:
    case REW_SMOOTH_FAST_CONVERGENCE:
    case REW_SMOOTHER_FAST_CONVERGENCE: {
        double nr = inf_norm(res, M);
        double extra = 1;
        if (converged) { double k = (double)(cfg->max_iters + 1 - steps); extra = k * k * 10; }
        double rew = (nr == 0) ? 1000 : -log(nr);
        if (cfg->strategy == REW_SMOOTH_FAST_CONVERGENCE && rew > 1) rew = 1 + log(rew);
        rew *= extra;
        if (cfg->strategy == REW_SMOOTHER_FAST_CONVERGENCE && rew > 1) rew = 1 + log(rew);
        return rew;
    }
    default:
        if (domain_err) *domain_err = 2;
        return NAN;
    }
}

/* reset: u = 1, r = u0 - C@u (sdc_env.py:306-314), for N lambdas */
void sdc_oracle_reset(int M, const double *Q, double dt, int64_t N, const double *lam, double *u_, double *r_,
                      int variant) {
    cplx C[MAXM * MAXM];
    for (int64_t e = 0; e < N; e++) {
        cplx z = { lam[2 * e] * dt, lam[2 * e + 1] * dt };
        cplx *u = (cplx *)u_ + e * M, *r = (cplx *)r_ + e * M;
        build_C(M, Q, z, C);
        for (int i = 0; i < M; i++) { u[i].re = 1.0; u[i].im = 0.0; }
        for (int i = 0; i < M; i++) {
            cplx y = zgemv_rowdot(C + i * M, u, M, variant);
            r[i].re = 1.0 - y.re;
            r[i].im = 0.0 - y.im;
        }
    }
}

static void get_scaled_action(int A, const double *action, int action_is_complex, int do_scale, int64_t e, cplx *d) {
    for (int k = 0; k < A; k++) {
        if (action_is_complex) {
            d[k].re = action[(e * A + k) * 2];
            d[k].im = action[(e * A + k) * 2 + 1];
        } else {
            double a = action[e * A + k];
            d[k].re = (do_scale & 1) ? scale_action(a) : a;
            d[k].im = 0.0;
        }
    }
}

/*
 * sdc-v1: one sweep per step (sdc_env.py:507-572).  In/out: u, r, niter.  Out per env: reward, done (0/1),
 * resnorm (info['residual']), err (0/1).  rinit = initial residual of the episode (for residual_change).
 * Pinv_out (optional, [N][M][M][2]) receives the emulated inverse for inspection.
 */
void sdc_oracle_step_v1(int M, const double *Q, double dt, int64_t N, int prec_type, const double *Qd_fixed,
                        const double *action, int action_is_complex, int do_scale, const double *lam,
                        double *u_, double *r_, int32_t *niter, const double *rinit_,
                        int strategy, double step_penalty, double residual_weight, double norm_factor,
                        double restol, int max_iters,
                        double *reward, uint8_t *done, double *resnorm, uint8_t *err_out, int variant,
                        double *Pinv_out) {
    reward_cfg cfg = { strategy, step_penalty, residual_weight, norm_factor, restol, max_iters };
    int A = sdc_oracle_num_actions(M, prec_type);
    cplx C[MAXM * MAXM], Qd[MAXM * MAXM], P[MAXM * MAXM], Pinv[MAXM * MAXM], d[MAXM * MAXM], old_r[MAXM];
    for (int64_t e = 0; e < N; e++) {
        cplx z = { lam[2 * e] * dt, lam[2 * e + 1] * dt };
        cplx *u = (cplx *)u_ + e * M, *r = (cplx *)r_ + e * M;
        const cplx *rinit = (const cplx *)rinit_ + e * M;
        build_C(M, Q, z, C);
        get_scaled_action(A, action, action_is_complex, do_scale, e, d);
        build_Qd(M, prec_type, d, Qd_fixed, Qd);
        build_P(M, z, Qd, P, (do_scale & 2) && prec_type != PREC_FIXED);
        sdc_oracle_cinv(M, (const double *)P, (double *)Pinv, variant);
        if (Pinv_out) memcpy(Pinv_out + e * M * M * 2, Pinv, sizeof(cplx) * M * M);
        memcpy(old_r, r, sizeof(cplx) * M);
        for (int i = 0; i < M; i++) {
            cplx dl = zgemv_rowdot(Pinv + i * M, old_r, M, variant);
            u[i] = cadd(u[i], dl);
        }
        for (int i = 0; i < M; i++) {
            cplx y = zgemv_rowdot(C + i * M, u, M, variant);
            r[i].re = 1.0 - y.re;
            r[i].im = 0.0 - y.im;
        }
        double nr = inf_norm(r, M), nr_old = inf_norm(old_r, M);
        niter[e] += 1;
        int err = isnan(nr) || isinf(nr);
        err = err || nr > nr_old * 100;
        int dn = nr < restol;
        double rew;
        if (!err) rew = reward_func(&cfg, M, old_r, r, rinit, dn, niter[e], 0);
        else rew = -step_penalty * (max_iters + 1);
        dn = dn || niter[e] >= max_iters || err;
        reward[e] = rew;
        done[e] = (uint8_t)dn;
        resnorm[e] = nr;
        if (err_out) err_out[e] = (uint8_t)err;
    }
}

/*
 * sdc-v0: full solve per step (sdc_env.py:209-273).  niter is an output (reset to 0 at loop start).
 * old_states (optional, [N][2M][max_iters] complex): column `niter` written when niter < max_iters (:239-240).
 */
void sdc_oracle_step_v0(int M, const double *Q, double dt, int64_t N, int prec_type, const double *Qd_fixed,
                        const double *action, int action_is_complex, int do_scale, const double *lam,
                        double *u_, double *r_, int32_t *niter, const double *rinit_,
                        int strategy, double step_penalty, double residual_weight, double norm_factor,
                        double restol, int max_iters,
                        double *reward, uint8_t *converged_out, double *resnorm, uint8_t *err_out, int variant,
                        double *old_states) {
    reward_cfg cfg = { strategy, step_penalty, residual_weight, norm_factor, restol, max_iters };
    int A = sdc_oracle_num_actions(M, prec_type);
    cplx C[MAXM * MAXM], Qd[MAXM * MAXM], P[MAXM * MAXM], Pinv[MAXM * MAXM], d[MAXM * MAXM], dl[MAXM];
    for (int64_t e = 0; e < N; e++) {
        cplx z = { lam[2 * e] * dt, lam[2 * e + 1] * dt };
        cplx *u = (cplx *)u_ + e * M, *r = (cplx *)r_ + e * M;
        const cplx *rinit = (const cplx *)rinit_ + e * M;
        build_C(M, Q, z, C);
        get_scaled_action(A, action, action_is_complex, do_scale, e, d);
        build_Qd(M, prec_type, d, Qd_fixed, Qd);
        build_P(M, z, Qd, P, (do_scale & 2) && prec_type != PREC_FIXED);
        sdc_oracle_cinv(M, (const double *)P, (double *)Pinv, variant);
        double nr_old = inf_norm(r, M), nr = nr_old;
        int dn = 0, err = 0, it = 0;
        double rew = 0;
        while (!dn && !(it >= max_iters)) {
            it++;
            for (int i = 0; i < M; i++) dl[i] = zgemv_rowdot(Pinv + i * M, r, M, variant);
            for (int i = 0; i < M; i++) u[i] = cadd(u[i], dl[i]);
            for (int i = 0; i < M; i++) {
                cplx y = zgemv_rowdot(C + i * M, u, M, variant);
                r[i].re = 1.0 - y.re;
                r[i].im = 0.0 - y.im;
            }
            nr = inf_norm(r, M);
            err = isnan(nr) || isinf(nr);
            if (old_states && it < max_iters) {
                cplx *os = (cplx *)old_states + e * (2 * M) * max_iters;
                for (int i = 0; i < M; i++) { os[i * max_iters + it] = u[i]; os[(M + i) * max_iters + it] = r[i]; }
            }
            err = err || nr > nr_old * 100;
            if (err) { rew = -step_penalty * (max_iters + 1); break; }
            dn = nr < restol;
        }
        if (!err) rew = reward_func(&cfg, M, rinit, r, rinit, dn, it, 0);
        niter[e] = it;
        reward[e] = rew;
        if (converged_out) converged_out[e] = (uint8_t)dn;
        resnorm[e] = nr;
        if (err_out) err_out[e] = (uint8_t)err;
    }
}
