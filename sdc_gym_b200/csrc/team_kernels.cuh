// team_kernels.cuh - step kernel for dense Q_delta at large M (6..9): one env per TEAM of M lanes, lane = node.
//
// With one env per thread the M x M inverse (2 M^2 doubles), its LU work matrix and C no longer fit the register
// file for M >= 6 (they spill to local memory, or shared memory leaves 2-4 warps per SM).  Here lane `row` of a
// team owns row `row` of every matrix (row of P/LU, row of the inverse, row of C) and its own u, r component, so
// the per-lane state is O(M) and everything stays in registers.  Vectors are exchanged with warp shuffles by
// absolute lane index (teams need not be a power of two wide: 32/M envs per warp, the few left-over lanes shadow
// rows of the last team and store nothing).
//
// The arithmetic is the same rounding sequence as the per-thread kernels (exact_math.cuh / exact_inv_reg.cuh,
// SURVEY Appendix A / A.2): every row of zgetf2 / ztrsm / zgemv_t is an independent chain, so distributing rows
// over lanes changes who computes a value, never how.  Parity: tests/test_gpu_parity.py (golden vectors + oracle).
#pragma once
#include "step_kernels.cuh"

#ifdef __CUDACC__
namespace sdcgym {

constexpr int kTeamBlock = 128;
#ifndef SDCGYM_TEAM_MIN_M
#define SDCGYM_TEAM_MIN_M 8
#endif
constexpr int kTeamMinM = SDCGYM_TEAM_MIN_M;
#ifndef SDCGYM_TEAM_MINB
#define SDCGYM_TEAM_MINB 3
#endif
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

template <int M>
__device__ __forceinline__ void gather(double v, int base, double (&out)[M]) {
#pragma unroll
    for (int j = 0; j < M; j++) out[j] = shfl_d(v, base + j);
}

// ---- np.linalg.inv for one env per team.
//   zgetf2: lane `prow` owns one row of the work matrix.  Row interchanges never move data: `holder[i]` is the
//           lane (team-relative) that holds logical row i, `myrow` the logical row this lane holds, and every
//           "row i" of the algorithm is addressed through holder[] (shuffle source) / myrow (predicate).
//   zgetrs: the M right-hand sides are independent, lane `prow` solves column `prow` of B = P I with the per-thread
//           ztrsm sequence; every LU entry is broadcast once from the lane that owns its row.
//   The inverse then sits column-per-lane; it is transposed through shared memory (`tp`, M*M complex per team) so
//   that lane `prow` ends up with row `prow` of Pinv for the sweeps.
//   In: ar/ai = row `prow` of P (destroyed).  Out: br/bi = row `prow` of inv(P). ----
template <int M, int V>
__device__ __forceinline__ void team_cinv(int prow, int base, bool writer, double (&ar)[M], double (&ai)[M],
                                          double2* tp, double (&br)[M], double (&bi)[M]) {
    int myrow = prow;
    int holder[M];
#pragma unroll
    for (int i = 0; i < M; i++) holder[i] = i;
    // ------------------------------- zgetf2 (left-looking) -------------------------------
#pragma unroll
    for (int j = 0; j < M; j++) {
        // ztrsv_NLU on the column head: for i < j: rows k in (i, j): b_k += (-b_i) * A(k, i)
#pragma unroll
        for (int i = 0; i < j; i++) {
            const cplx alpha{-shfl_d(ar[j], base + holder[i]), -shfl_d(ai[j], base + holder[i])};
            if (myrow > i && myrow < j) {
                const cplx pr = cmul_blas<V>(alpha, cplx{ar[i], ai[i]});
                ar[j] = dadd(pr.re, ar[j]);
                ai[j] = dadd(pr.im, ai[j]);
            }
        }
        // zgemv_n: rows >= j: b -= A[row, 0:j] @ b[0:j]
        if (j > 0) {
            double xr[M], xi[M];  // b[0:j] (final after the trsv)
#pragma unroll
            for (int c = 0; c < M; c++) {
                if (c < j) {
                    xr[c] = shfl_d(ar[j], base + holder[c]);
                    xi[c] = shfl_d(ai[j], base + holder[c]);
                }
            }
            const int rows = M - j, r4 = rows & ~3;
            const int ii = myrow - j;
            if (ii >= 0 && ii < r4) {
                double ybr = 0.0, ybi = 0.0;
#pragma unroll
                for (int c0 = 0; c0 < M; c0++) {
                    int w = 0;
                    if (c0 < (j & ~3)) w = ((c0 & 3) == 0) ? 4 : 0;
                    else if (c0 == (j & ~3) && (j & 2)) w = 2;
                    else if (c0 == (j & ~1) && (j & 1)) w = 1;
                    if (w > 0) {
                        double S1 = dmul(xr[c0], ar[c0]), S2 = dmul(xr[c0], ai[c0]);
                        double S3 = dmul(xi[c0], ar[c0]), S4 = dmul(xi[c0], ai[c0]);
#pragma unroll
                        for (int q = 1; q < 4; q++) {
                            if (q < w) {
                                const int c = c0 + q;
                                if (c < M) {
                                    S1 = dfma(xr[c], ar[c], S1);
                                    S2 = dfma(xr[c], ai[c], S2);
                                    S3 = dfma(xi[c], ar[c], S3);
                                    S4 = dfma(xi[c], ai[c], S4);
                                }
                            }
                        }
                        ybr = dadd(ybr, dsub(S1, S4));
                        ybi = dadd(ybi, dadd(S2, S3));
                    }
                }
                ar[j] = dadd(ar[j], -ybr);
                ai[j] = dadd(ai[j], -ybi);
            } else if (ii >= r4 && ii < rows) {
                double tr = 0.0, ti = 0.0;
#pragma unroll
                for (int c = 0; c < M; c++) {
                    if (c < j) {
                        const cplx pr = cmul_blas<V>(cplx{ar[c], ai[c]}, cplx{xr[c], xi[c]});
                        tr = dadd(tr, pr.re);
                        ti = dadd(ti, pr.im);
                    }
                }
                ar[j] = dadd(-tr, ar[j]);
                ai[j] = dadd(-ti, ai[j]);
            }
        }
        // pivot: first logical row >= j maximising |re| + |im|
        const double mine = dadd(fabs(ar[j]), fabs(ai[j]));
        int p = j;
        double best = shfl_d(mine, base + holder[j]);
#pragma unroll
        for (int i = j + 1; i < M; i++) {
            const double v = shfl_d(mine, base + holder[i]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        // interchange logical rows j <-> p: bookkeeping only
#pragma unroll
        for (int q = j + 1; q < M; q++) {
            const bool sw = (p == q);
            const int t = holder[j];
            holder[j] = sw ? holder[q] : t;
            holder[q] = sw ? t : holder[q];
        }
        myrow = (myrow == j) ? p : ((myrow == p) ? j : myrow);
        // scale the sub-column by the (unfused) pivot reciprocal with unfused products (zscal)
        const cplx inv = crecip<false>(cplx{shfl_d(ar[j], base + holder[j]), shfl_d(ai[j], base + holder[j])});
        if (myrow > j) {
            const cplx s = cmul_unfused(inv, cplx{ar[j], ai[j]});
            ar[j] = s.re;
            ai[j] = s.im;
        }
    }

    // ------------------------------- zgetrs: column `prow` of B = P I, two ztrsm -------------------------------
    using T = TrsmTiles<M>;
    // reciprocal of the diagonal entry of the logical row this lane holds (only the upper solve uses it)
    double dr = 0.0, di = 0.0;
#pragma unroll
    for (int c = 0; c < M; c++)
        if (c == myrow) {
            dr = ar[c];
            di = ai[c];
        }
    const cplx invd = crecip<V == 0>(cplx{dr, di});
    // entry (i, pp) of the LU factors, broadcast from the lane that holds logical row i
    auto A = [&](int i, int pp) { return cplx{shfl_d(ar[pp], base + holder[i]), shfl_d(ai[pp], base + holder[i])}; };

    double bcr[M], bci[M];
#pragma unroll
    for (int i = 0; i < M; i++) {
        bcr[i] = (holder[i] == prow) ? 1.0 : 0.0;  // row i of P I is e_{perm[i]}, perm[i] = the lane row i started on
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
                                    const cplx a = A(i, pp);
                                    Srr = dfma(a.re, bcr[pp], Srr);
                                    Sii = dfma(a.im, bci[pp], Sii);
                                    Sri = dfma(a.re, bci[pp], Sri);
                                    Sir = dfma(a.im, bcr[pp], Sir);
                                }
                            }
                            vr = dsub(Srr, Sii);
                            vi = dadd(Sir, Sri);
                        } else {
                            double re = 0.0, im = 0.0;
#pragma unroll
                            for (int pp = 0; pp < M; pp++) {
                                if (pp >= p_lo && pp < p_hi) {
                                    const cplx a = A(i, pp);
                                    re = dfma(bcr[pp], a.re, -dfma(bci[pp], a.im, -re));
                                    im = dfma(bcr[pp], a.im, dfma(bci[pp], a.re, im));
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
                    if (upper) {
                        const cplx d{shfl_d(invd.re, base + holder[i]), shfl_d(invd.im, base + holder[i])};
                        ccv = cmul_blas<V>(d, ccv);
                    }
                    bcr[i] = ccv.re;
                    bci[i] = ccv.im;
#pragma unroll
                    for (int s2 = 1; s2 < 4; s2++) {
                        if (s2 > s && s2 < rs) {
                            const int k = upper ? r0 + rs - 1 - s2 : r0 + s2;
                            const cplx pr = cmul_blas<V>(ccv, A(k, i));
                            bcr[k] = dsub(bcr[k], pr.re);
                            bci[k] = dsub(bci[k], pr.im);
                        }
                    }
                }
            }
        }
    }
    // transpose: lane `prow` holds column `prow`; it needs row `prow`
#pragma unroll
    for (int i = 0; i < M; i++)
        if (writer) tp[i * M + prow] = make_double2(bcr[i], bci[i]);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < M; c++) {
        const double2 v = tp[prow * M + c];
        br[c] = v.x;
        bi[c] = v.y;
    }
    __syncwarp();
}

// =====================================================================================================
// step kernel, team layout (dense Q_delta only)
// =====================================================================================================
template <int M, int KIND, int V, int MINB = SDCGYM_TEAM_MINB>
__global__ void __launch_bounds__(kTeamBlock, MINB) team_step_kernel(const __grid_constant__ StepParams<M> p) {
    constexpr int TPW = 32 / M;  // envs per warp
    constexpr int WPB = kTeamBlock / 32;
    // per team: M*M complex for the transpose of the inverse; afterwards its first 2*M slots are the exchange
    // buffers of the sweeps (u at [0, M), r at [M, 2M))
    __shared__ double2 team_smem[WPB * TPW * M * M];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * WPB + warp;
    const bool lane_used = lane < TPW * M;
    const int team = lane_used ? lane / M : TPW - 1;
    const int row = lane_used ? lane - team * M : (lane - TPW * M) % M;  // left-over lanes shadow rows of the last team
    const int base = team * M;
    double2* const tsm = team_smem + (warp * TPW + team) * (M * M);
    const int64_t env = warp_global * TPW + team;
    const bool valid = lane_used && env < p.N;
    const int64_t i = env < p.N ? env : p.N - 1;
    const int64_t ld = p.ld;

    const double lr = p.lam[i], li = p.lam[ld + i];
    const double zr = dmul(lr, p.dt), zi = dmul(li, p.dt);

    // ---- state: own component (loaded first: the latency hides behind the inverse) ----
    double ur = p.S[(2 * row) * ld + i], ui = p.S[(2 * row + 1) * ld + i];
    double rr = p.S[(2 * M + 2 * row) * ld + i], ri = p.S[(2 * M + 2 * row + 1) * ld + i];
    const double nr_old = p.resnorm[i];
    int it = (KIND == SDCGYM_ENV_STEP) ? p.niter[i] : 0;
    // read now what lane `row == 0` rewrites at the end (every lane of the team needs the old values)
    const int32_t ep_old = p.autoreset ? p.episodes[i] : 0;
    const uint32_t ctr_old = p.autoreset ? p.rng_ctr[i] : 0u;

    // ---- row `row` of P = eye(M) - (lam*dt)*Qd ----
    double ar[M], ai[M], pr_[M], pi_[M];
#pragma unroll
    for (int c = 0; c < M; c++) {
        cplx d{0.0, 0.0};
        int k = -1;
        switch (p.prec_type) {
        case SDCGYM_PREC_LOWER_DIAG: k = (row == c + 1) ? c : -1; break;
        case SDCGYM_PREC_LOWER_TRI: k = (c <= row) ? row * (row + 1) / 2 + c : -1; break;
        case SDCGYM_PREC_STRICTLY_LOWER_TRI: k = (c < row) ? row * (row - 1) / 2 + c : -1; break;
        case SDCGYM_PREC_DIAG: k = (c == row) ? row : -1; break;
        default: break;
        }
        if (p.prec_type == SDCGYM_PREC_FIXED) {
            d.re = p.Qd[row * M + c];
        } else if (k >= 0) {
            if (p.is_complex) {
                d.re = ld_ro(p.action + i * p.a_es + k * p.a_cs);
                d.im = ld_ro(p.action + i * p.a_es + k * p.a_cs + 1);
            } else {
                const double a = ld_ro(p.action + i * p.a_es + k * p.a_cs);
                d.re = a;
                if (p.do_scale & SDCGYM_ACTION_SCALE) d.re = (a <= -1.0) ? 0.0 : ((a >= 1.0) ? 1.0 : dmul(0.5, dadd(a, 1.0)));
            }
        }
        const cplx zq = ((p.do_scale & SDCGYM_ACTION_F32) && p.prec_type != SDCGYM_PREC_FIXED)
                            ? cmul_np_f32(cplx{zr, zi}, d) : cmul_np(cplx{zr, zi}, d);
        ar[c] = dsub((row == c) ? 1.0 : 0.0, zq.re);
        ai[c] = dsub(0.0, zq.im);
    }
    team_cinv<M, V>(row, base, lane_used, ar, ai, tsm, pr_, pi_);  // pr_/pi_ = row of Pinv

    // ---- row `row` of C = eye(M) - (lam*dt)*Q  (reuses ar/ai) ----
#pragma unroll
    for (int c = 0; c < M; c++) {
        const double q = p.Q[row * M + c];
        ar[c] = (row == c) ? dsub(1.0, dmul(zr, q)) : -dmul(zr, q);
        ai[c] = -dmul(zi, q);
    }


    // exchange of a team-wide complex vector: every lane publishes its component, then reads all M
    double2* const xu = tsm;
    double2* const xr = tsm + M;
    auto exchange = [&](double2* buf, double re, double im, double (&Vr)[M], double (&Vi)[M]) {
        if (lane_used) buf[row] = make_double2(re, im);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; m++) {
            const double2 v = buf[m];
            Vr[m] = v.x;
            Vi[m] = v.y;
        }
    };

    double Rr[M], Ri[M];  // the full residual vector, identical in every lane of the team
    exchange(xr, rr, ri, Rr, Ri);
    double norm_old_scaled = nr_old;
    if (KIND == SDCGYM_ENV_STEP && p.strategy == SDCGYM_REW_RESIDUAL_CHANGE && p.norm_factor != 1.0)
        norm_old_scaled = scaled_inf_norm<M>(Rr, Ri, p.norm_factor);

    // one sweep: u += Pinv @ r ; r = u0 - C @ u.  Teams that are finished (`act` false) run along (the exchanges are
    // warp-wide) but commit nothing.
    auto sweep = [&](bool act) {
        double dr, di, yr, yi, Ur[M], Ui[M];
        zgemv_rowdot<M, V>(pr_, pi_, Rr, Ri, dr, di);
        const double nur = dadd(ur, dr), nui = dadd(ui, di);
        ur = act ? nur : ur;
        ui = act ? nui : ui;
        exchange(xu, ur, ui, Ur, Ui);
        zgemv_rowdot<M, V>(ar, ai, Ur, Ui, yr, yi);
        rr = act ? dsub(1.0, yr) : rr;
        ri = act ? -yi : ri;
        exchange(xr, rr, ri, Rr, Ri);
    };

    const double thr = dmul(nr_old, 100.0);
    bool conv = false, err = false;
    double nr = nr_old;

    if (KIND == SDCGYM_ENV_FULL) {
        const HiBand bc = make_band(p.restol), be = make_band(thr);
        const SqBand sc = make_sqband(p.restol), se = make_sqband(thr);
        bool act = p.max_iters > 0;
        while (__any_sync(kFull, act)) {
            sweep(act);
            if (act) {
                it++;
                const int H = absmax_hi<M>(Rr, Ri);
                const bool amb_e = (H > be.lo) && (H < be.hi), amb_c = (H > bc.lo) && (H < bc.hi);
                if (H >= be.hi) {
                    err = true;
                } else if (!(amb_e || amb_c)) {
                    conv = H <= bc.lo;
                } else {
                    const double s2 = sq_absmax<M>(Rr, Ri);
                    const int ge = !amb_e ? 0 : (!se.ok ? 2 : (s2 < se.lo2 ? 0 : (s2 > se.hi2 ? 1 : 2)));
                    const int gc = !amb_c ? (H <= bc.lo ? 0 : 1) : (!sc.ok ? 2 : (s2 < sc.lo2 ? 0 : (s2 > sc.hi2 ? 1 : 2)));
                    if (ge == 2 || gc == 2) {
                        double tr[M], ti[M];
#pragma unroll
                        for (int m = 0; m < M; m++) {
                            tr[m] = Rr[m];
                            ti[m] = Ri[m];
                        }
                        const double nx = inf_norm_slow<M>(tr, ti);
                        err = isnan(nx) || isinf(nx) || nx > thr;
                        if (!err) conv = nx < p.restol;
                    } else {
                        err = ge == 1;
                        conv = !err && gc == 0;
                    }
                }
                act = !err && !conv && it < p.max_iters;
            }
        }
        if (p.max_iters > 0) nr = inf_norm_fast<M>(Rr, Ri);
    } else {
        sweep(true);
        nr = inf_norm_fast<M>(Rr, Ri);
        it++;
        err = isnan(nr) || isinf(nr);
        err = err || nr > thr;
        conv = nr < p.restol;
    }

    // ---- reward (every lane of the team computes the same value; lane `row == 0` stores it) ----
    double rew;
    if (err) {
        rew = dmul(-p.step_penalty, (double)(p.max_iters + 1));
    } else if (p.strategy == SDCGYM_REW_ITERATION_ONLY) {
        rew = dmul((double)(-it), p.step_penalty);
    } else {
        double norm_init_scaled = 0.0;
        if (p.strategy == SDCGYM_REW_RESIDUAL_CHANGE) {
            double tu[M], tv[M], ir[M], ii[M];
            initial_state<M, V>(p.Q, zr, zi, tu, tv, ir, ii);
            norm_init_scaled = scaled_inf_norm<M>(ir, ii, p.norm_factor);
            if (KIND == SDCGYM_ENV_FULL) norm_old_scaled = norm_init_scaled;
        }
        double tr[M], ti[M];  // copies: the callee is not inlined and takes the arrays by reference
#pragma unroll
        for (int m = 0; m < M; m++) {
            tr[m] = Rr[m];
            ti[m] = Ri[m];
        }
        rew = reward_func<M>(p.strategy, p.step_penalty, p.residual_weight, p.norm_factor, p.restol, p.max_iters,
                             norm_old_scaled, norm_init_scaled, tr, ti, nr, conv, it, p.log_restol_nf);
    }
    const bool done = (KIND == SDCGYM_ENV_FULL) ? true : (conv || it >= p.max_iters || err);

    // ---- DummyVecEnv auto-reset: every lane derives the team's new lambda and its own component of the new
    //      residual r = u0 - C @ 1; the team-wide vector (for the new ||r||inf) goes through the exchange buffer.
    //      Computed for every team (the exchange is warp-wide), used by the ones that are done. ----
    double nlr = 0.0, nli = 0.0, nrr = 0.0, nri = 0.0, nres = 0.0, ninit = 0.0;
    const int32_t ep = ep_old + 1;
    if (p.autoreset) {
        draw_lambda<M>(p, i, ctr_old, ep, nlr, nli);
        const double nzr = dmul(nlr, p.dt), nzi = dmul(nli, p.dt);
        double cr[M], ci[M], one[M], zero[M], yr, yi, Ir[M], Ii[M];
#pragma unroll
        for (int c = 0; c < M; c++) {
            const double q = p.Q[row * M + c];
            cr[c] = (row == c) ? dsub(1.0, dmul(nzr, q)) : -dmul(nzr, q);
            ci[c] = -dmul(nzi, q);
            one[c] = 1.0;
            zero[c] = 0.0;
        }
        zgemv_rowdot<M, V>(cr, ci, one, zero, yr, yi);
        nrr = dsub(1.0, yr);
        nri = -yi;
        __syncwarp();  // every lane has read the last residual exchange
        exchange(xr, nrr, nri, Ir, Ii);
        nres = inf_norm_fast<M>(Ir, Ii);
        ninit = (KIND == SDCGYM_ENV_STEP && p.norm_init && p.norm_factor != 1.0) ? scaled_inf_norm<M>(Ir, Ii, p.norm_factor) : nres;
    }
    if (!valid) return;

    if (row == 0) {
        if (p.reward) p.reward[i] = rew;
        if (p.flags)
            p.flags[i] = (uint8_t)((done ? SDCGYM_FLAG_DONE : 0) | (conv ? SDCGYM_FLAG_CONVERGED : 0) | (err ? SDCGYM_FLAG_ERR : 0));
        if (p.info_res) p.info_res[i] = nr;
        if (p.info_niter) p.info_niter[i] = it;
        if (p.info_lam) {
            p.info_lam[2 * i] = lr;
            p.info_lam[2 * i + 1] = li;
        }
    }
    auto store_own = [&](double* __restrict__ S, double a, double b, double c, double d) {
        S[(2 * row) * ld + i] = a;
        S[(2 * row + 1) * ld + i] = b;
        S[(2 * M + 2 * row) * ld + i] = c;
        S[(2 * M + 2 * row + 1) * ld + i] = d;
    };
    if (done && p.term) store_own(p.term, ur, ui, rr, ri);

    if (done && p.autoreset) {
        store_own(p.S, 1.0, 0.0, nrr, nri);
        if (row == 0) {
            p.episodes[i] = ep;
            p.rng_ctr[i] = ctr_old + 1;
            p.lam[i] = nlr;
            p.lam[ld + i] = nli;
            p.resnorm[i] = nres;
            if (KIND == SDCGYM_ENV_STEP && p.norm_init) p.norm_init[i] = ninit;
            p.niter[i] = 0;
        }
    } else {
        store_own(p.S, ur, ui, rr, ri);
        if (row == 0) {
            p.resnorm[i] = nr;
            p.niter[i] = it;
        }
    }
}

}  // namespace sdcgym
#endif  // __CUDACC__
