// fast_kernels.cuh - the CERTIFIED SUBSTITUTION sweep mode of the sdc-v0 full solve (SDCGYM_SWEEP_CERTIFIED).
//
// Reference statement (sdc_env.py:229-231; node-by-node form in sdc_env_nonlinear.py:248-264):
//     (I - z Q_delta) u+ = u0 + z (Q - Q_delta) u      <=>      u+ = u + Pinv r,   r+ = u0 - u+ + z (Q u+)
// The exact mode (step_kernels.cuh) reproduces the reference's numpy/OpenBLAS rounding sequence: a dense complex
// `C @ u` per sweep (183 FP64 instructions at M = 5).  Here the residual is formed from the REAL collocation matrix
// (Q u: 2 M^2 FMAs with constant-bank operands, then one complex scale by z) and the update is two FMAs per component:
// 95 FP64 instructions per M = 5 sweep, nothing per-env but u, r, Pinv, z in registers, no shared memory.
//
// The rounding sequence differs from the reference's, so every decision the reference takes on ||r||inf
// (`nr > 100 nr_old` -> err, `nr < restol` -> done; sdc_env.py:241-247) is taken here WITH A MARGIN that bounds
// | ||r~_k|| - ||r^ref_k|| | rigorously (certify.cuh: explicit powers of the iteration matrix give ||K^n|| <= G theta^n;
// the kernel carries F_{k+1} = theta F_k + eta_k, E_k = G F_k >= ||u~_k - u^ref_k||_w and
// margin_k = Gamma E_k + b0 + b1 U_k).  A decision outside the margin is provably the reference's.  An env that meets
// a decision inside the margin writes NOTHING and puts its index on the fallback list; the exact kernel re-runs those
// envs from their untouched state (step_list_kernel).  Results: niter / done / converged / err bit-equal to the
// reference for every env, u, r, ||r||, reward of the certified envs within rounding (<= 1e-12 relative; the north
// star's tolerance), those of the fallback envs bit-equal.
#pragma once
#include "certify.cuh"
#include "step_kernels.cuh"

namespace sdcgym {

struct FastWork {
    float* cert;      // [kCertPlanes][ld]
    int32_t* list;    // [N] indices of envs that need the exact kernel
    int32_t* count;   // [0] = length of `list` for the running step, [1] = cumulative over steps
};

SDCGYM_HD double from_hi(int h) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(h, 0);
#else
    uint64_t u = (uint64_t)(uint32_t)h << 32;
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

// P = I - z*Qd diagonal and its reciprocals: the very statements of the exact kernel (step_one prologue), so Pinv has
// the same bits in both modes
template <int M, int V>
SDCGYM_HD void diag_pinv(const StepParams<M>& p, double zr, double zi, const double (&araw)[M], const double (&aimg)[M],
                         double (&Pr)[M], double (&Pi)[M]) {
    double pre_r[M], pre_i[M];
#pragma unroll
    for (int k = 0; k < M; k++) {
        cplx zq;
        if (p.prec_type == SDCGYM_PREC_FIXED) {
            const double d = p.Qd[k * M + k];
            zq = cplx{dmul(zr, d), dmul(zi, d)};
        } else if (p.is_complex) {
            cplx d{araw[k], aimg[k]};
            zq = (p.do_scale & SDCGYM_ACTION_F32) ? cmul_np_f32(cplx{zr, zi}, d) : cmul_np(cplx{zr, zi}, d);
        } else {
            double a = araw[k];
            double d = a;
            if (p.do_scale & SDCGYM_ACTION_SCALE) d = (a <= -1.0) ? 0.0 : ((a >= 1.0) ? 1.0 : dmul(0.5, dadd(a, 1.0)));
            zq = (p.do_scale & SDCGYM_ACTION_F32) ? cmul_np_f32(cplx{zr, zi}, cplx{d, 0.0}) : cplx{dmul(zr, d), dmul(zi, d)};
        }
        pre_r[k] = dsub(1.0, zq.re);
        pre_i[k] = -zq.im;
    }
    crecip_batch<M, V == 0>(pre_r, pre_i, Pr, Pi);
}

template <int M>
SDCGYM_HD void load_actions(const StepParams<M>& p, int64_t i, double (&araw)[M], double (&aimg)[M]) {
#pragma unroll
    for (int k = 0; k < M; k++) {
        araw[k] = 0.0;
        aimg[k] = 0.0;
        if (p.prec_type != SDCGYM_PREC_FIXED) {
            araw[k] = ld_ro(p.action + i * p.a_es + k * p.a_cs);
            if (p.is_complex) aimg[k] = ld_ro(p.action + i * p.a_es + k * p.a_cs + 1);
        }
    }
}

// ---- certificate kernel body: one env ----
template <int M, int V>
SDCGYM_HD void cert_one(const StepParams<M>& p, const FastWork& fw, int64_t i) {
    if (i >= p.N) return;
    const double lr = p.lam[i], li = p.lam[p.ld + i];
    double araw[M], aimg[M];
    load_actions<M>(p, i, araw, aimg);
    const double zr = dmul(lr, p.dt), zi = dmul(li, p.dt);
    double Pr[M], Pi[M];
    diag_pinv<M, V>(p, zr, zi, araw, aimg, Pr, Pi);
    const Cert c = cert_diag<M, kCertPowers>(p.Q, zr, zi, Pr, Pi);
    float* o = fw.cert + i;
    const int64_t ld = p.ld;
    o[0 * ld] = c.theta;
    o[1 * ld] = c.G;
    o[2 * ld] = c.gam;
    o[3 * ld] = c.a0;
    o[4 * ld] = c.a1;
    o[5 * ld] = c.a2;
    o[6 * ld] = c.b0;
    o[7 * ld] = c.b1;
}

// max_j |u_j| <= max_j |Re u_j| + max_j |Im u_j|: two integer-pipe maxima of high words, rounded up to the next high word.
// (Tighter than sqrt(2) max(|Re|, |Im|) where it matters: good preconditioners keep u close to the real axis.)
template <int M>
SDCGYM_HD double modulus_bound(const double (&vr)[M], const double (&vi)[M]) {
    int hr = 0, hi = 0;
#pragma unroll
    for (int m = 0; m < M; m++) {
        hr = imax(hr, hi_word(vr[m]) & 0x7fffffff);
        hi = imax(hi, hi_word(vi[m]) & 0x7fffffff);
    }
    return dadd(from_hi(hr + 1), from_hi(hi + 1));
}

// three-way comparison of the unknown reference norm nr^ref in [lo - margin, up + margin] with a threshold t
//   +1: nr^ref > t for sure     -1: nr^ref < t for sure     0: undecided
SDCGYM_HD int sure_cmp(double lo, double up, double margin, double t) {
    if (lo - margin > t) return 1;
    if (up + margin < t) return -1;
    return 0;
}

// second-stage decisions (rare): bits 0-1 = sure_cmp against thr + 1, bits 2-3 = sure_cmp against restol + 1
template <int M>
SDCGYM_HD_NOINLINE int stage2_decide(const double (&rr)[M], const double (&ri)[M], double margin, double thr, double restol) {
    const double s2 = sq_absmax<M>(rr, ri);
    if (!(s2 > 1e-280 && s2 < 1e300)) return 1 | (1 << 2);  // undecided
    const double nr2 = dsqrt(s2);
    const double lo = dmul(nr2, 1.0 - 2e-15), up = dmul(nr2, 1.0 + 2e-15);
    return (sure_cmp(lo, up, margin, thr) + 1) | ((sure_cmp(lo, up, margin, restol) + 1) << 2);
}

// ---- the substitution step: one env per thread.  TRACE (host tests only): trace[k*(2M+2) ...] = r~_k (2M), ||r~_k||, margin_k ----
template <int M, int V, bool TRACE = false>
SDCGYM_HD void fast_step_one(const StepParams<M>& p, const FastWork& fw, const int64_t tid, double* trace = nullptr) {
    const bool valid = tid < p.N;
    const int64_t i = valid ? tid : p.N - 1;
    const int64_t ld = p.ld;
    constexpr double kSqrt2Up = 1.41421356237310;  // > sqrt(2)

    // ---- every global load up front (one DRAM round trip) ----
    double lr = p.lam[i], li = p.lam[ld + i];
    double araw[M], aimg[M];
    load_actions<M>(p, i, araw, aimg);
    double ur[M], ui[M], rr[M], ri[M];
#pragma unroll
    for (int m = 0; m < M; m++) {
        ur[m] = p.S[(2 * m) * ld + i];
        ui[m] = p.S[(2 * m + 1) * ld + i];
        rr[m] = p.S[(2 * M + 2 * m) * ld + i];
        ri[m] = p.S[(2 * M + 2 * m + 1) * ld + i];
    }
    const double nr_old = p.resnorm[i];
    const int32_t ep_old = p.autoreset ? p.episodes[i] : 0;
    const uint32_t ctr_old = p.autoreset ? p.rng_ctr[i] : 0u;
    const float* cp = fw.cert + i;
    const double theta = (double)cp[0 * ld], G = (double)cp[1 * ld], gam = (double)cp[2 * ld];
    const double a0 = (double)cp[3 * ld], a1 = (double)cp[4 * ld], a2 = (double)cp[5 * ld];
    const double b0 = (double)cp[6 * ld], b1 = (double)cp[7 * ld];

    const double zr = dmul(lr, p.dt), zi = dmul(li, p.dt);
    double Pr[M], Pi[M];
    diag_pinv<M, V>(p, zr, zi, araw, aimg, Pr, Pi);

    const double thr = dmul(nr_old, 100.0);  // norm_res_old * 100 (same bits as the reference: the start state is exact)
    const double restol = p.restol;
    bool conv = false, err = false;
    bool amb = !(theta < 1e30) || !(G < 1e30);  // no usable certificate
    int it = 0;
    double F = 0.0, E = 0.0, margin = 0.0;
    double Ub = modulus_bound<M>(ur, ui);  // >= max_j |u~_j|  (Inf / NaN components give Inf / NaN: ambiguous below)
    int Hr = absmax_hi<M>(rr, ri);
    bool act = p.max_iters > 0 && !amb;
    while (SDCGYM_WARP_ANY(act)) {
        if (act) {
            it++;
            // local error of this sweep (certify.cuh): eta <= a0 + a1 (U_k + E_k) + a2 (R_k + margin_k)
            const double Ud = dadd(Ub, E);
            const double Rd = dfma(from_hi(Hr + 1), kSqrt2Up, margin);
            const double eta = dfma(a1, Ud, dfma(a2, Rd, a0));
            F = dfma(theta, F, eta);
            E = dmul(G, F);
            // u <- u + Pinv r
#pragma unroll
            for (int m = 0; m < M; m++) {
                const double nur = dfma(Pr[m], rr[m], dfma(-Pi[m], ri[m], ur[m]));
                const double nui = dfma(Pr[m], ri[m], dfma(Pi[m], rr[m], ui[m]));
                ur[m] = nur;
                ui[m] = nui;
            }
            Ub = modulus_bound<M>(ur, ui);
            // r = u0 - u + z (Q u)
#pragma unroll
            for (int m = 0; m < M; m++) {
                double sr = dmul(p.Q[m * M], ur[0]), si = dmul(p.Q[m * M], ui[0]);
#pragma unroll
                for (int j = 1; j < M; j++) {
                    sr = dfma(p.Q[m * M + j], ur[j], sr);
                    si = dfma(p.Q[m * M + j], ui[j], si);
                }
                rr[m] = dfma(zr, sr, dfma(-zi, si, dsub(1.0, ur[m])));
                ri[m] = dfma(zr, si, dfma(zi, sr, -ui[m]));
            }
            Hr = absmax_hi<M>(rr, ri);
            // decision margin for ||r_{k+1}||: Gamma E + b0 + b1 (U + E), plus the rounding of the reference's own norm
            double lo = from_hi(Hr), up = dmul(from_hi(Hr + 1), kSqrt2Up);
            const double Un = dadd(Ub, E);
            margin = dfma(gam, E, dfma(b1, Un, b0));
            margin = dfma(up, 2e-15, margin);
            if (Hr >= 0x7ff00000 || !(margin < 1e300)) {
                amb = true;  // Inf / NaN in the substitution residual, or the bound overflowed: the exact kernel decides
            } else {
                int ce = sure_cmp(lo, up, margin, thr), cc = sure_cmp(lo, up, margin, restol);
                if (ce == 0 || (ce < 0 && cc == 0)) {
                    // second stage: the norm itself (squared magnitudes: relative error < 1e-15).  Out of line (and on a
                    // private copy, so that rr / ri stay in registers): inlined, the compiler speculates its dozen
                    // FP64 instructions into every sweep.
                    double tr[M], ti[M];
#pragma unroll
                    for (int m = 0; m < M; m++) {
                        tr[m] = rr[m];
                        ti[m] = ri[m];
                    }
                    const int both = stage2_decide<M>(tr, ti, margin, thr, restol);
                    ce = (both & 3) - 1;
                    cc = (both >> 2) - 1;
                }
                if (ce > 0) err = true;
                else if (ce == 0) amb = true;
                else if (cc < 0) conv = true;
                else if (cc == 0) amb = true;
            }
            if (TRACE && trace && valid) {
                double* t = trace + (size_t)(it - 1) * (2 * M + 2);
                for (int m = 0; m < M; m++) {
                    t[2 * m] = rr[m];
                    t[2 * m + 1] = ri[m];
                }
                t[2 * M] = dsqrt(sq_absmax<M>(rr, ri));
                t[2 * M + 1] = margin;
            }
            act = !amb && !err && !conv && it < p.max_iters;
        }
    }

    // ---- ambiguous: leave the env untouched and hand it to the exact kernel ----
#ifdef __CUDA_ARCH__
    {
        const unsigned mask = __ballot_sync(0xffffffffu, amb && valid);
        if (mask) {
            const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(fw.count, __popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (amb && valid) fw.list[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)i;
        }
    }
#else
    if (amb && valid) fw.list[fw.count[0]++] = (int32_t)i;
#endif
    if (amb || !valid) return;

    double nr = nr_old;
    if (p.max_iters > 0) nr = dsqrt(sq_absmax<M>(rr, ri));

    // ---- reward (sdc_env.py:242-257), as in the exact kernel ----
    double rew;
    if (err) {
        rew = dmul(-p.step_penalty, (double)(p.max_iters + 1));
    } else if (p.strategy == SDCGYM_REW_ITERATION_ONLY) {
        rew = dmul((double)(-it), p.step_penalty);
    } else {
        double norm_init_scaled = 0.0, norm_old_scaled = nr_old;
        if (p.strategy == SDCGYM_REW_RESIDUAL_CHANGE) {
            double tu[M], tv[M], ir[M], ii[M];
            initial_state<M, V>(p.Q, zr, zi, tu, tv, ir, ii);
            norm_init_scaled = scaled_inf_norm<M>(ir, ii, p.norm_factor);
            norm_old_scaled = norm_init_scaled;
        }
        rew = reward_func<M>(p.strategy, p.step_penalty, p.residual_weight, p.norm_factor, p.restol, p.max_iters,
                             norm_old_scaled, norm_init_scaled, rr, ri, nr, conv, it, p.log_restol_nf);
    }

    if (p.reward) p.reward[i] = rew;
    if (p.flags) p.flags[i] = (uint8_t)(SDCGYM_FLAG_DONE | (conv ? SDCGYM_FLAG_CONVERGED : 0) | (err ? SDCGYM_FLAG_ERR : 0));
    if (p.info_res) p.info_res[i] = nr;
    if (p.info_niter) p.info_niter[i] = it;
    if (p.info_lam) {
        p.info_lam[2 * i] = lr;
        p.info_lam[2 * i + 1] = li;
    }
    if (p.term) store_state<M>(p.term, ld, i, ur, ui, rr, ri);

    if (p.autoreset) {
        // DummyVecEnv: obs = env.reset() right after the terminal step - exact arithmetic, it is the next start state
        const int32_t ep = ep_old + 1;
        p.episodes[i] = ep;
        double nlr, nli;
        draw_lambda<M>(p, i, ctr_old, ep, nlr, nli);
        p.rng_ctr[i] = ctr_old + 1;
        p.lam[i] = nlr;
        p.lam[ld + i] = nli;
        const double nzr = dmul(nlr, p.dt), nzi = dmul(nli, p.dt);
        initial_state<M, V>(p.Q, nzr, nzi, ur, ui, rr, ri);
        store_state<M>(p.S, ld, i, ur, ui, rr, ri);
        const double n0 = inf_norm_fast<M>(rr, ri);
        p.resnorm[i] = n0;  // (no norm_init store: sdc-v0 kernels do not keep that plane, see step_one)
        p.niter[i] = 0;
    } else {
        store_state<M>(p.S, ld, i, ur, ui, rr, ri);
        p.resnorm[i] = nr;
        p.niter[i] = it;
    }
}

#ifdef __CUDACC__
constexpr int kFastBlock = 128;

constexpr int kCertBlock = 128;
template <int M, int V>
__global__ void __launch_bounds__(kCertBlock, 3) cert_kernel(const __grid_constant__ StepParams<M> p, const FastWork fw) {
    const int64_t i = (int64_t)blockIdx.x * kCertBlock + threadIdx.x;
    if (i == 0) fw.count[0] = 0;  // the fast kernel (next launch on the stream) fills the list
    cert_one<M, V>(p, fw, i);
}

template <int M, int V, int MINB>
__global__ void __launch_bounds__(kFastBlock, MINB) fast_step_kernel(const __grid_constant__ StepParams<M> p, const FastWork fw) {
    fast_step_one<M, V>(p, fw, (int64_t)blockIdx.x * kFastBlock + threadIdx.x);
}

// the exact kernel over the fallback list (fixed grid, grid-stride over the list; lanes past the end run the clamped
// env and store nothing, exactly like the tail of a full launch)
template <int M, int V, int HOLD, int MINB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB) step_list_kernel(const __grid_constant__ StepParams<M> p, const FastWork fw) {
    const int count = fw.count[0];
    if (blockIdx.x == 0 && threadIdx.x == 0 && count > 0) atomicAdd(fw.count + 1, count);
    for (int base = blockIdx.x * BLOCK; base < count; base += gridDim.x * BLOCK) {
        const int t = base + threadIdx.x;
        const int64_t idx = (t < count) ? (int64_t)fw.list[t] : p.N;
        if constexpr (HOLD == 5 || HOLD == 7) {
            extern __shared__ double2 pside_smem[];
            step_one<M, SDCGYM_ENV_FULL, V, false, HOLD>(p, idx, nullptr, 1, reinterpret_cast<cplx*>(pside_smem) + threadIdx.x, BLOCK);
        } else if constexpr (HOLD >= 3) {
            extern __shared__ double side_smem[];
            step_one<M, SDCGYM_ENV_FULL, V, false, HOLD>(p, idx, side_smem + threadIdx.x, BLOCK);
        } else {
            step_one<M, SDCGYM_ENV_FULL, V, false, HOLD>(p, idx);
        }
    }
}
#endif

}  // namespace sdcgym
