// exact_inv_reg.cuh - register-resident version of cinv_exact (exact_math.cuh) for small M.
//
// Same arithmetic, rounding for rounding (np.linalg.inv = OpenBLAS zgetf2 + zgetrs), but
// every loop is unrolled over compile-time indices so the LU factors and the inverse live in registers instead
// of per-thread local memory (which made the dense-Q_delta kernels L2-bandwidth bound: ~1500 local accesses of
// 512 B per warp).  The data-dependent partial pivoting becomes a chain of predicated row swaps; zgetf2's lazy
// "apply earlier swaps to this column" is replaced by swapping whole rows eagerly, which moves the same values
// through the same sequence of swaps.
#pragma once
#include "exact_math.cuh"

namespace sdcgym {

template <int M>
struct TrsmTiles {
    // row tiles in storage order: 4,4,...,(2),(1)
    static constexpr int n4 = M / 4, has2 = (M & 2) ? 1 : 0, has1 = M & 1, nrt = n4 + has2 + has1;
    SDCGYM_HD static constexpr int start(int t) { return t < n4 ? 4 * t : (has2 && t == n4 ? 4 * n4 : 4 * n4 + 2 * has2); }
    SDCGYM_HD static constexpr int size(int t) { return t < n4 ? 4 : (has2 && t == n4 ? 2 : 1); }
    // visiting order of the backward (upper) solve: 1-tile, 2-tile, then 4-tiles bottom to top
    SDCGYM_HD static constexpr int upper_visit(int oi) {
        return oi < has1 ? nrt - 1 : (oi < has1 + has2 ? n4 : n4 - 1 - (oi - has1 - has2));
    }
};

// storage of the LU work matrix: registers (two arrays) or a strided side store (shared memory, element k of this
// thread at base[k * stride]; static indices either way)
template <int M>
struct RegMatrix {
    double re[M * M], im[M * M];
    SDCGYM_HD double& R(int i, int j) { return re[i + j * M]; }
    SDCGYM_HD double& I(int i, int j) { return im[i + j * M]; }
};
template <int M>
struct StridedMatrix {
    // volatile: every use is a load from the side store; a cached copy would be a register (and then a spill) again
    volatile double* base;
    int stride;
    SDCGYM_HD volatile double& R(int i, int j) { return base[(2 * (i + j * M)) * stride]; }
    SDCGYM_HD volatile double& I(int i, int j) { return base[(2 * (i + j * M) + 1) * stride]; }
};

#define AR_(i, j) A.R((i), (j))
#define AI_(i, j) A.I((i), (j))

// in: A = P (column-major, split re/im), destroyed (holds LU afterwards).  The inverse is produced one column at a
// time (the two triangular solves of different columns are independent) and handed to store(row, col, re, im), so
// only A and one column of B are live: M = 6, 7 still fit the register file this way.
template <int M, int V, class Mat, class Store>
SDCGYM_HD void cinv_exact_reg_cols(Mat& A, Store&& store) {
    int perm[M];
#pragma unroll
    for (int i = 0; i < M; i++) perm[i] = i;

    // ------------------------------- zgetf2 (left-looking) -------------------------------
#pragma unroll
    for (int j = 0; j < M; j++) {
        // ztrsv_NLU on the column head b[0:j]
#pragma unroll
        for (int i = 0; i < j; i++) {
            const cplx alpha{-AR_(i, j), -AI_(i, j)};
#pragma unroll
            for (int k = i + 1; k < j; k++) {
                const cplx pr = cmul_blas<V>(alpha, cplx{AR_(k, i), AI_(k, i)});
                AR_(k, j) = dadd(pr.re, AR_(k, j));
                AI_(k, j) = dadd(pr.im, AI_(k, j));
            }
        }
        // zgemv_n: b[j:] -= A[j:, 0:j] @ b[0:j]
        if (j > 0) {
            constexpr int dummy = 0;
            (void)dummy;
            const int rows = M - j, r4 = rows & ~3;
#pragma unroll
            for (int ii = 0; ii < M; ii++) {
                if (ii < r4) {
                    const int i = j + ii;
                    double ybr = 0.0, ybi = 0.0;
                    // column blocks: 4,4,..., then 2 if (j & 2), then 1 if (j & 1)
#pragma unroll
                    for (int c0 = 0; c0 < M; c0++) {
                        // block starts: multiples of 4 below (j & ~3), then (j & ~3) [if j&2], then (j & ~1) [if j&1]
                        int w = 0;
                        if (c0 < (j & ~3)) w = ((c0 & 3) == 0) ? 4 : 0;
                        else if (c0 == (j & ~3) && (j & 2)) w = 2;
                        else if (c0 == (j & ~1) && (j & 1)) w = 1;
                        if (w > 0) {
                            double S1 = dmul(AR_(c0, j), AR_(i, c0)), S2 = dmul(AR_(c0, j), AI_(i, c0));
                            double S3 = dmul(AI_(c0, j), AR_(i, c0)), S4 = dmul(AI_(c0, j), AI_(i, c0));
#pragma unroll
                            for (int q = 1; q < 4; q++) {
                                if (q < w) {
                                    const int c = c0 + q;
                                    if (c < M) {
                                        S1 = dfma(AR_(c, j), AR_(i, c), S1);
                                        S2 = dfma(AR_(c, j), AI_(i, c), S2);
                                        S3 = dfma(AI_(c, j), AR_(i, c), S3);
                                        S4 = dfma(AI_(c, j), AI_(i, c), S4);
                                    }
                                }
                            }
                            ybr = dadd(ybr, dsub(S1, S4));
                            ybi = dadd(ybi, dadd(S2, S3));
                        }
                    }
                    AR_(i, j) = dadd(AR_(i, j), -ybr);
                    AI_(i, j) = dadd(AI_(i, j), -ybi);
                } else if (ii < rows) {
                    const int i = j + ii;
                    double tr = 0.0, ti = 0.0;
#pragma unroll
                    for (int c = 0; c < j; c++) {
                        const cplx pr = cmul_blas<V>(cplx{AR_(i, c), AI_(i, c)}, cplx{AR_(c, j), AI_(c, j)});
                        tr = dadd(tr, pr.re);
                        ti = dadd(ti, pr.im);
                    }
                    AR_(i, j) = dadd(-tr, AR_(i, j));
                    AI_(i, j) = dadd(-ti, AI_(i, j));
                }
            }
        }
        // pivot: first row >= j maximising |re| + |im|
        int p = j;
        double best = dadd(fabs(AR_(j, j)), fabs(AI_(j, j)));
#pragma unroll
        for (int i = j + 1; i < M; i++) {
            const double v = dadd(fabs(AR_(i, j)), fabs(AI_(i, j)));
            if (v > best) {
                best = v;
                p = i;
            }
        }
        // eager swap of rows j <-> p over all columns (predicated), and of the permutation vector
#pragma unroll
        for (int q = j + 1; q < M; q++) {
            const bool sw = (p == q);
#pragma unroll
            for (int c = 0; c < M; c++) {
                const double tr = AR_(j, c), ti = AI_(j, c);
                AR_(j, c) = sw ? AR_(q, c) : tr;
                AI_(j, c) = sw ? AI_(q, c) : ti;
                AR_(q, c) = sw ? tr : AR_(q, c);
                AI_(q, c) = sw ? ti : AI_(q, c);
            }
            const int tp = perm[j];
            perm[j] = sw ? perm[q] : tp;
            perm[q] = sw ? tp : perm[q];
        }
        // scale the sub-column by the (unfused) pivot reciprocal with unfused products (zscal)
        const cplx inv = crecip<false>(cplx{AR_(j, j), AI_(j, j)});
#pragma unroll
        for (int k = j + 1; k < M; k++) {
            const cplx s = cmul_unfused(inv, cplx{AR_(k, j), AI_(k, j)});
            AR_(k, j) = s.re;
            AI_(k, j) = s.im;
        }
    }

    // ------------------------------- zgetrs: B = P I, then two ztrsm, column by column -------------------------------
    using T = TrsmTiles<M>;
    double invr[M], invi[M];
#pragma unroll
    for (int i = 0; i < M; i++) {
        const cplx d = crecip<V == 0>(cplx{AR_(i, i), AI_(i, i)});
        invr[i] = d.re;
        invi[i] = d.im;
    }
#pragma unroll
    for (int col = 0; col < M; col++) {
        double bcr[M], bci[M];
#pragma unroll
        for (int i = 0; i < M; i++) {
            bcr[i] = (perm[i] == col) ? 1.0 : 0.0;
            bci[i] = 0.0;
        }
#pragma unroll
        for (int upper = 0; upper < 2; upper++) {
#pragma unroll
            for (int oi = 0; oi < T::nrt; oi++) {
                const int t = upper ? T::upper_visit(oi) : oi;
                const int r0 = T::start(t), rs = T::size(t);
                const int p_lo = upper ? r0 + rs : 0, p_hi = upper ? M : r0;
                // (a) update with all already-solved rows
                if (p_hi > p_lo) {
#pragma unroll
                    for (int ii = 0; ii < 4; ii++) {
                        if (ii < rs) {
                            const int i = r0 + ii;
                            double vr, vi;
                            if (rs == 4) {
                                double Srr = 0.0, Sii = 0.0, Sri = 0.0, Sir = 0.0;
#pragma unroll
                                for (int pp = 0; pp < M; pp++) {
                                    if (pp >= p_lo && pp < p_hi) {
                                        Srr = dfma(AR_(i, pp), bcr[pp], Srr);
                                        Sii = dfma(AI_(i, pp), bci[pp], Sii);
                                        Sri = dfma(AR_(i, pp), bci[pp], Sri);
                                        Sir = dfma(AI_(i, pp), bcr[pp], Sir);
                                    }
                                }
                                vr = dsub(Srr, Sii);
                                vi = dadd(Sir, Sri);
                            } else {
                                double re = 0.0, im = 0.0;
#pragma unroll
                                for (int pp = 0; pp < M; pp++) {
                                    if (pp >= p_lo && pp < p_hi) {
                                        re = dfma(bcr[pp], AR_(i, pp), -dfma(bci[pp], AI_(i, pp), -re));
                                        im = dfma(bcr[pp], AI_(i, pp), dfma(bci[pp], AR_(i, pp), im));
                                    }
                                }
                                vr = re;
                                vi = im;
                            }
                            bcr[i] = dsub(bcr[i], vr);
                            bci[i] = dsub(bci[i], vi);
                        }
                    }
                }
                // (b) in-tile solve (ascending rows forward, descending backward)
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    if (s < rs) {
                        const int i = upper ? r0 + rs - 1 - s : r0 + s;
                        cplx ccv{bcr[i], bci[i]};
                        if (upper) ccv = cmul_blas<V>(cplx{invr[i], invi[i]}, ccv);
                        bcr[i] = ccv.re;
                        bci[i] = ccv.im;
#pragma unroll
                        for (int s2 = 1; s2 < 4; s2++) {
                            if (s2 > s && s2 < rs) {
                                const int k = upper ? r0 + rs - 1 - s2 : r0 + s2;
                                const cplx pr = cmul_blas<V>(ccv, cplx{AR_(k, i), AI_(k, i)});
                                bcr[k] = dsub(bcr[k], pr.re);
                                bci[k] = dsub(bci[k], pr.im);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < M; i++) store(i, col, bcr[i], bci[i]);
    }
}

// convenience form: inverse into column-major arrays B
template <int M, int V>
SDCGYM_HD void cinv_exact_reg(double (&Ar)[M * M], double (&Ai)[M * M], double (&Br)[M * M], double (&Bi)[M * M]) {
    RegMatrix<M> A;
#pragma unroll
    for (int k = 0; k < M * M; k++) {
        A.re[k] = Ar[k];
        A.im[k] = Ai[k];
    }
    cinv_exact_reg_cols<M, V>(A, [&](int i, int c, double re, double im) {
        Br[i + c * M] = re;
        Bi[i + c * M] = im;
    });
}

#undef AR_
#undef AI_

}  // namespace sdcgym
