// exact_math.cuh - fp64 primitives that reproduce, rounding for rounding, the arithmetic the reference env
// executes through numpy / OpenBLAS (oracle/sdc_exact.c).
//
// Device build: every operation is spelled with an explicit round-to-nearest intrinsic (__dmul_rn,
// __dadd_rn, __fma_rn, __ddiv_rn, __dsqrt_rn) so nvcc can neither contract a*b+c into an FMA nor split
// one; the translation units are additionally compiled with -fmad=false.
// Host build (tests/host_shim only - lets the CPU-only test suite execute the very same templates that the
// kernels instantiate; never part of the product library): plain IEEE operators under -ffp-contract=off
// and std::fma.
//
// V (blas variant): 0 = OpenBLAS "SkylakeX" core (scalar C tails contracted to FMA), 1 = "Haswell" core.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <math_constants.h>
#define SDCGYM_HD __host__ __device__ __forceinline__
#define SDCGYM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SDCGYM_HD inline
#define SDCGYM_HD_NOINLINE inline
#endif

namespace sdcgym {

#ifdef __CUDA_ARCH__
SDCGYM_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
SDCGYM_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
SDCGYM_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
SDCGYM_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
SDCGYM_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
SDCGYM_HD double dsqrt(double a) { return __dsqrt_rn(a); }
SDCGYM_HD int hi_word(double x) { return __double2hiint(x); }
SDCGYM_HD double ld_ro(const double* p) { return __ldg(p); }
#else
SDCGYM_HD double dmul(double a, double b) { return a * b; }
SDCGYM_HD double dadd(double a, double b) { return a + b; }
SDCGYM_HD double dsub(double a, double b) { return a - b; }
SDCGYM_HD double dfma(double a, double b, double c) { return fma(a, b, c); }
SDCGYM_HD double ddiv(double a, double b) { return a / b; }
SDCGYM_HD double dsqrt(double a) { return sqrt(a); }
SDCGYM_HD int hi_word(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int)(u >> 32);
}
SDCGYM_HD double ld_ro(const double* p) { return *p; }
#endif
SDCGYM_HD double d_inf() { return (double)INFINITY; }
SDCGYM_HD double d_nan() { return (double)NAN; }

struct cplx {
    double re, im;
};

// numpy complex multiply loop (reference `lam * dt * Q`, sdc_env.py:199,304): fused form on every build
// probed (Appendix A step 0): (fms(ar,br, ai*bi), fma(ar,bi, ai*br)).
SDCGYM_HD cplx cmul_np(cplx a, cplx b) {
    cplx c;
    c.re = dfma(a.re, b.re, -dmul(a.im, b.im));
    c.im = dfma(a.re, b.im, dmul(a.im, b.re));
    return c;
}
// use_doubles=False (float32 / complex64 action space, sdc_env.py:100,109): Qdmat is allocated in the action dtype
// (:138-140), so `fill_diagonal` rounds the scaled action to float32, and `self.lam * self.dt * Qdmat` (:199) - a
// Python complex (lam is built from Python floats, :293-300) times a float32/complex64 array - is evaluated by numpy
// in complex64: operands rounded to float32, same fused loop form as above in float arithmetic.  `eye(M) - ...`
// then promotes the product to complex128 exactly.
SDCGYM_HD cplx cmul_np_f32(cplx a, cplx b) {
    const float ar = (float)a.re, ai = (float)a.im, br = (float)b.re, bi = (float)b.im;
    cplx c;
#ifdef __CUDA_ARCH__
    c.re = (double)__fmaf_rn(ar, br, -__fmul_rn(ai, bi));
    c.im = (double)__fmaf_rn(ar, bi, __fmul_rn(ai, br));
#else
    c.re = (double)fmaf(ar, br, -(ai * bi));
    c.im = (double)fmaf(ar, bi, ai * br);
#endif
    return c;
}
SDCGYM_HD cplx cmul_unfused(cplx a, cplx b) {
    cplx c;
    c.re = dsub(dmul(a.re, b.re), dmul(a.im, b.im));
    c.im = dadd(dmul(a.re, b.im), dmul(a.im, b.re));
    return c;
}
// complex product as compiled inside OpenBLAS' scalar C kernels (contracted on SkylakeX only)
template <int V>
SDCGYM_HD cplx cmul_blas(cplx a, cplx b) {
    if (V == 0) return cmul_np(a, b);
    return cmul_unfused(a, b);
}

// OpenBLAS complex reciprocal (ztrsm diagonal inverse / zgetf2 pivot), Appendix A step 6.
template <bool FUSED>
SDCGYM_HD cplx crecip(cplx p) {
    // branch-free form of
    //   |pr| >= |pi| : t = pi/pr, den = 1/(pr*(1+t*t)), inv = (den, -t*den)
    //   else         : t = pr/pi, den = 1/(pi*(1+t*t)), inv = (t*den, -den)
    // (same operations on selected operands; -t*den == -(t*den) exactly), so independent reciprocals interleave.
    const bool first = fabs(p.re) >= fabs(p.im);
    const double a = first ? p.re : p.im, b = first ? p.im : p.re;
    const double t = ddiv(b, a);
    const double tt = FUSED ? dfma(t, t, 1.0) : dadd(1.0, dmul(t, t));
    const double den = ddiv(1.0, dmul(a, tt));
    const double td = dmul(t, den);
    cplx inv;
    inv.re = first ? den : td;
    inv.im = first ? -td : -den;
    return inv;
}

// numpy |z| (AVX512F loop): L*sqrt(fma(s/L, s/L, 1)), L = max(|re|,|im|), s = min (Appendix A step 9).
SDCGYM_HD double np_cabs(double re, double im) {
    double a = fabs(re), b = fabs(im);
    if (isinf(a) || isinf(b)) return d_inf();
    if (isnan(a) || isnan(b)) return d_nan();
    double L = a > b ? a : b, s = a > b ? b : a;
    if (L == 0.0) return 0.0;
    double t = ddiv(s, L);
    return dmul(L, dsqrt(dfma(t, t, 1.0)));
}

// np.linalg.norm(v, inf) = abs(v).max(), NaN-propagating (sdc_env.py:206-207)
template <int M>
SDCGYM_HD double inf_norm(const double (&vr)[M], const double (&vi)[M]) {
    double m = -d_inf();
    bool nan = false;
#pragma unroll
    for (int i = 0; i < M; i++) {
        double a = np_cabs(vr[i], vi[i]);
        nan |= isnan(a);
        m = a > m ? a : m;
    }
    return nan ? d_nan() : m;
}

// max over all 2M components of the high word of |x|: a monotone 32-bit proxy of max(|re|,|im|) that runs
// entirely on the integer pipe.  from_hi(H) <= max|x| < from_hi(H+1); Inf/NaN give H >= 0x7ff00000.
SDCGYM_HD int imax(int a, int b) { return a > b ? a : b; }

template <int M>
SDCGYM_HD int absmax_hi(const double (&vr)[M], const double (&vi)[M]) {
    int h = 0;
#pragma unroll
    for (int i = 0; i < M; i++) {
        h = imax(h, hi_word(vr[i]) & 0x7fffffff);
        h = imax(h, hi_word(vi[i]) & 0x7fffffff);
    }
    return h;
}

// One row of numpy `A @ x` for a C-contiguous complex128 (M,M) matrix: OpenBLAS zgemv_t micro-kernel
// (Appendix A step 3).  a = matrix row, x = vector.  The leading `0.0 +` of the BLAS kernel only affects
// the sign of an exact zero and is dropped.
template <int M, int V>
SDCGYM_HD void zgemv_rowdot(const double (&ar)[M], const double (&ai)[M], const double (&xr)[M],
                                             const double (&xi)[M], double& yr, double& yi) {
    constexpr int m3 = M & 3, m1 = M - m3;
    yr = 0.0;
    yi = 0.0;
    if (m1 > 0) {
        double E1 = dmul(xr[0], ar[0]), E2 = dmul(xr[0], ai[0]), E3 = dmul(xi[0], ar[0]), E4 = dmul(xi[0], ai[0]);
        double O1 = dmul(xr[1], ar[1]), O2 = dmul(xr[1], ai[1]), O3 = dmul(xi[1], ar[1]), O4 = dmul(xi[1], ai[1]);
#pragma unroll
        for (int j = 2; j < m1; j += 2) {
            E1 = dfma(xr[j], ar[j], E1);
            E2 = dfma(xr[j], ai[j], E2);
            E3 = dfma(xi[j], ar[j], E3);
            E4 = dfma(xi[j], ai[j], E4);
            O1 = dfma(xr[j + 1], ar[j + 1], O1);
            O2 = dfma(xr[j + 1], ai[j + 1], O2);
            O3 = dfma(xi[j + 1], ar[j + 1], O3);
            O4 = dfma(xi[j + 1], ai[j + 1], O4);
        }
        yr = dadd(dsub(O1, O4), dsub(E1, E4));
        yi = dadd(dadd(O2, O3), dadd(E2, E3));
    }
    if (m3 > 0) {
        cplx t = cmul_blas<V>(cplx{ar[m1], ai[m1]}, cplx{xr[m1], xi[m1]});
#pragma unroll
        for (int j = m1 + 1; j < M; j++) {
            cplx e = cmul_blas<V>(cplx{ar[j], ai[j]}, cplx{xr[j], xi[j]});
            t.re = dadd(e.re, t.re);
            t.im = dadd(e.im, t.im);
        }
        if (m1 > 0) {
            yr = dadd(t.re, yr);
            yi = dadd(t.im, yi);
        } else {
            yr = t.re;  // t + (+0): value-identical
            yi = t.im;
        }
    }
}

// `Pinv @ r` when Pinv is exactly diagonal: the zgemv_t dot of row i degenerates to a single complex
// product whose rounding pattern depends on whether column i sits in the vector head (i < m1: four plain
// products) or in the scalar tail (i >= m1: fused on SkylakeX).  Appendix A step 7.  a = Pinv_ii, x = r_i.
template <int M, int V>
SDCGYM_HD void diag_rowdot(int i, double ar, double ai, double xr, double xi, double& yr, double& yi) {
    constexpr int m1 = M - (M & 3);
    if (i < m1) {
        yr = dsub(dmul(xr, ar), dmul(xi, ai));
        yi = dadd(dmul(xr, ai), dmul(xi, ar));
    } else {
        cplx e = cmul_blas<V>(cplx{ar, ai}, cplx{xr, xi});
        yr = e.re;
        yi = e.im;
    }
}

// Second-stage classification on squared magnitudes: s_i = fma(re, re, im*im) carries a relative error below
// 3e-16 and numpy's |v_i| below 5e-16, so  max_i s_i < t^2 (1 - 1e-14)  proves ||v||inf < t and
// max_i s_i > t^2 (1 + 1e-14)  proves ||v||inf > t.  Only the sliver in between (probability ~1e-14 per
// decision) needs the exact div + sqrt norm.  Costs 2 FP64 instructions per node instead of ~70.
struct SqBand {
    double lo2, hi2;
    bool ok;  // t^2 representable without over/underflow
};
SDCGYM_HD SqBand make_sqband(double t) {
    SqBand b;
    b.ok = (t > 1e-140) && (t < 1e140);
    const double t2 = dmul(t, t);
    b.lo2 = dmul(t2, 1.0 - 1e-14);
    b.hi2 = dmul(t2, 1.0 + 1e-14);
    return b;
}
template <int M>
SDCGYM_HD double sq_absmax(const double (&vr)[M], const double (&vi)[M]) {
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        const double sq = dfma(vr[k], vr[k], dmul(vi[k], vi[k]));
        m = sq > m ? sq : m;
    }
    return m;
}

// out-of-line exact norm for the rare paths (keeps the sweep loop small)
template <int M>
SDCGYM_HD_NOINLINE double inf_norm_slow(const double* vr, const double* vi) {
    double m = -d_inf();
    bool nan = false;
    for (int k = 0; k < M; k++) {
        double a = np_cabs(vr[k], vi[k]);
        nan |= isnan(a);
        m = a > m ? a : m;
    }
    return nan ? d_nan() : m;
}

// np.linalg.norm(v, inf), bit-exact, evaluating the div + sqrt of numpy's |.| only for the entry that can
// attain the maximum: if s_k is the only squared magnitude within 1e-14 of the largest one, every other
// |v_j| is provably smaller than |v_k| (see SqBand) and the norm is |v_k|.
template <int M>
SDCGYM_HD double inf_norm_fast(const double (&vr)[M], const double (&vi)[M]) {
    double sq[M];
    double smax = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        sq[k] = dfma(vr[k], vr[k], dmul(vi[k], vi[k]));
        smax = sq[k] > smax ? sq[k] : smax;
    }
    bool fine = (smax > 1e-280) && (smax < 1e300);
    const double cut = dmul(smax, 1.0 - 1e-14);
    int cand = 0;
    double br = 0.0, bi = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        fine = fine && (sq[k] == sq[k]);  // no NaN component
        if (sq[k] >= cut) {
            cand++;
            br = vr[k];
            bi = vi[k];
        }
    }
    if (fine && cand == 1) return np_cabs(br, bi);
    double tr[M], ti[M];  // private copy: keeps the caller's arrays in registers
#pragma unroll
    for (int k = 0; k < M; k++) {
        tr[k] = vr[k];
        ti[k] = vi[k];
    }
    return inf_norm_slow<M>(tr, ti);
}

// ---------------------------------------------------------------------------------------------------
// np.linalg.inv for M <= 9 = LAPACK zgesv(P, I): OpenBLAS zgetf2 (left-looking, partial pivoting) then
// zgetrs = row swaps + ztrsm(unit lower) + ztrsm(upper).  Appendix A.2; mirrors oracle/sdc_exact.c
// zgetf2_emul / ztrsm_emul.  Works on per-thread local arrays (column-major), data-dependent pivoting.
// ---------------------------------------------------------------------------------------------------
template <int M, int V>
SDCGYM_HD_NOINLINE void cinv_exact(cplx* __restrict__ A /*M*M col-major, in: P, out: LU*/,
                                        cplx* __restrict__ B /*M*M col-major, out: inverse*/,
                                        const int stride = 1 /*distance between consecutive elements, in cplx*/) {
    int ipiv[M];
    cplx b[M];
#define AT_(X, i, j) X[((i) + (j) * M) * stride]
#pragma unroll 1
    for (int j = 0; j < M; j++) {
        for (int i = 0; i < M; i++) b[i] = AT_(A, i, j);
        for (int i = 0; i < j; i++) {
            int p = ipiv[i];
            if (p != i) {
                cplx t = b[i];
                b[i] = b[p];
                b[p] = t;
            }
        }
        // ztrsv_NLU
        for (int i = 0; i < j; i++) {
            cplx alpha = cplx{-b[i].re, -b[i].im};
            for (int k = i + 1; k < j; k++) {
                cplx pr = cmul_blas<V>(alpha, AT_(A, k, i));
                b[k].re = dadd(pr.re, b[k].re);
                b[k].im = dadd(pr.im, b[k].im);
            }
        }
        // zgemv_n update of b[j:]
        if (j > 0) {
            int rows = M - j, r4 = rows & ~3;
            for (int ii = 0; ii < r4; ii++) {
                int i = j + ii;
                double ybr = 0.0, ybi = 0.0;
                int c = 0;
                int nblk4 = j >> 2;
                int nblk = nblk4 + ((j & 2) ? 1 : 0) + ((j & 1) ? 1 : 0);
                for (int blk = 0; blk < nblk; blk++) {
                    int w = blk < nblk4 ? 4 : ((blk == nblk4 && (j & 2)) ? 2 : 1);
                    double S1 = 0, S2 = 0, S3 = 0, S4 = 0;
                    for (int q = 0; q < w; q++, c++) {
                        cplx a = AT_(A, i, c), x = b[c];
                        if (q == 0) {
                            S1 = dmul(x.re, a.re);
                            S2 = dmul(x.re, a.im);
                            S3 = dmul(x.im, a.re);
                            S4 = dmul(x.im, a.im);
                        } else {
                            S1 = dfma(x.re, a.re, S1);
                            S2 = dfma(x.re, a.im, S2);
                            S3 = dfma(x.im, a.re, S3);
                            S4 = dfma(x.im, a.im, S4);
                        }
                    }
                    ybr = dadd(ybr, dsub(S1, S4));
                    ybi = dadd(ybi, dadd(S2, S3));
                }
                b[i].re = dadd(b[i].re, -ybr);
                b[i].im = dadd(b[i].im, -ybi);
            }
            for (int ii = r4; ii < rows; ii++) {
                int i = j + ii;
                double tr = 0.0, ti = 0.0;
                for (int c = 0; c < j; c++) {
                    cplx pr = cmul_blas<V>(AT_(A, i, c), b[c]);
                    tr = dadd(tr, pr.re);
                    ti = dadd(ti, pr.im);
                }
                b[i].re = dadd(-tr, b[i].re);
                b[i].im = dadd(-ti, b[i].im);
            }
        }
        // pivot: first row >= j maximising |re| + |im|
        int p = j;
        double best = dadd(fabs(b[j].re), fabs(b[j].im));
        for (int i = j + 1; i < M; i++) {
            double v = dadd(fabs(b[i].re), fabs(b[i].im));
            if (v > best) {
                best = v;
                p = i;
            }
        }
        ipiv[j] = p;
        for (int i = 0; i < M; i++) AT_(A, i, j) = b[i];
        if (p != j) {
            for (int c = 0; c <= j; c++) {
                cplx t = AT_(A, j, c);
                AT_(A, j, c) = AT_(A, p, c);
                AT_(A, p, c) = t;
            }
        }
        cplx inv = crecip<false>(AT_(A, j, j));
        for (int k = j + 1; k < M; k++) AT_(A, k, j) = cmul_unfused(inv, AT_(A, k, j));
    }

    // B = I with the row swaps applied
    for (int i = 0; i < M * M; i++) B[i * stride] = cplx{0.0, 0.0};
    for (int i = 0; i < M; i++) AT_(B, i, i).re = 1.0;
    for (int i = 0; i < M; i++) {
        int p = ipiv[i];
        if (p != i)
            for (int c = 0; c < M; c++) {
                cplx t = AT_(B, i, c);
                AT_(B, i, c) = AT_(B, p, c);
                AT_(B, p, c) = t;
            }
    }

    // two ztrsm passes: unit-lower forward, then non-unit upper backward.
    // Row tiles 4,4,..,(2),(1) in storage order; column tiles 2,2,..,(1).
    constexpr int n4 = M / 4, has2 = (M & 2) ? 1 : 0, has1 = M & 1;
    constexpr int nrt = n4 + has2 + has1;
    cplx invd[M];
    for (int i = 0; i < M; i++) invd[i] = crecip<V == 0>(AT_(A, i, i));
#pragma unroll 1
    for (int upper = 0; upper < 2; upper++) {
#pragma unroll 1
        for (int col0 = 0; col0 < M; col0 += 2) {
            int cw = (M - col0 >= 2) ? 2 : 1;
#pragma unroll 1
            for (int oi = 0; oi < nrt; oi++) {
                // tile index in storage order for the oi-th visit
                int t;
                if (!upper) {
                    t = oi;
                } else {
                    // remainder tiles first (1-tile, then 2-tile), then 4-tiles bottom to top
                    if (oi < has1) t = nrt - 1;
                    else if (oi < has1 + has2) t = n4;
                    else t = n4 - 1 - (oi - has1 - has2);
                }
                int r0, rs;
                if (t < n4) { r0 = 4 * t; rs = 4; }
                else if (has2 && t == n4) { r0 = 4 * n4; rs = 2; }
                else { r0 = 4 * n4 + 2 * has2; rs = 1; }
                for (int cc = 0; cc < cw; cc++) {
                    int col = col0 + cc;
                    int p_lo = upper ? r0 + rs : 0, p_hi = upper ? M : r0;
                    if (p_hi > p_lo) {
                        for (int ii = 0; ii < rs; ii++) {
                            int i = r0 + ii;
                            double vr, vi;
                            if (rs == 4) {
                                double Srr = 0, Sii = 0, Sri = 0, Sir = 0;
                                for (int p = p_lo; p < p_hi; p++) {
                                    cplx a = AT_(A, i, p), bb = AT_(B, p, col);
                                    Srr = dfma(a.re, bb.re, Srr);
                                    Sii = dfma(a.im, bb.im, Sii);
                                    Sri = dfma(a.re, bb.im, Sri);
                                    Sir = dfma(a.im, bb.re, Sir);
                                }
                                vr = dsub(Srr, Sii);
                                vi = dadd(Sir, Sri);
                            } else {
                                double re = 0, im = 0;
                                for (int p = p_lo; p < p_hi; p++) {
                                    cplx a = AT_(A, i, p), bb = AT_(B, p, col);
                                    re = dfma(bb.re, a.re, -dfma(bb.im, a.im, -re));
                                    im = dfma(bb.re, a.im, dfma(bb.im, a.re, im));
                                }
                                vr = re;
                                vi = im;
                            }
                            AT_(B, i, col).re = dsub(AT_(B, i, col).re, vr);
                            AT_(B, i, col).im = dsub(AT_(B, i, col).im, vi);
                        }
                    }
                    for (int s = 0; s < rs; s++) {
                        int i = upper ? r0 + rs - 1 - s : r0 + s;
                        cplx ccv = upper ? cmul_blas<V>(invd[i], AT_(B, i, col)) : AT_(B, i, col);
                        AT_(B, i, col) = ccv;
                        for (int s2 = s + 1; s2 < rs; s2++) {
                            int k = upper ? r0 + rs - 1 - s2 : r0 + s2;
                            cplx pr = cmul_blas<V>(ccv, AT_(A, k, i));
                            AT_(B, k, col).re = dsub(AT_(B, k, col).re, pr.re);
                            AT_(B, k, col).im = dsub(AT_(B, k, col).im, pr.im);
                        }
                    }
                }
            }
        }
    }
#undef AT_
}

}  // namespace sdcgym
