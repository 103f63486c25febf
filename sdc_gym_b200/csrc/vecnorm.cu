// vecnorm.cu - device-side VecNormalize (SB3 semantics restated: RunningMeanStd with Chan's merge, clip to
// +-clip) over the observation planes and the reward/return planes, plus the ResidualLoss sweep.
//
// Reference call sites: utils/utils.py:295-312 (VecNormalize(env, norm_obs, norm_reward, gamma)),
// dp_playground.py:235-258 (ResidualLoss.take_step).  SB3 itself is third-party and not installed: its
// documented semantics are restated ("parity unpinned").  Complex observations are normalised on
// their re / im planes separately (SB3 on complex128 is ill-defined).
//
// All reductions use a fixed block count and a fixed summation tree: deterministic run to run.
#include <cuda_runtime.h>
#include <string.h>
#include <math_constants.h>

#include "../../include/sdcgym.h"
#include "specrad.cuh"

namespace sdcgym {

#ifndef SDCGYM_ACC_BATCH
#define SDCGYM_ACC_BATCH 8
#endif
constexpr int kAccBlocks = 64, kAccThreads = 256, kAccBatch = SDCGYM_ACC_BATCH;  // (the batch only groups loads: the summation order is unchanged)

__device__ __forceinline__ double2 block_sum2(double a, double b) {
    __shared__ double sa[kAccThreads / 32], sb[kAccThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sa[threadIdx.x >> 5] = a;
        sb[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    a = (threadIdx.x < kAccThreads / 32) ? sa[threadIdx.x] : 0.0;
    b = (threadIdx.x < kAccThreads / 32) ? sb[threadIdx.x] : 0.0;
    if (threadIdx.x < 32)
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    __syncthreads();
    return make_double2(a, b);
}

// The statistics arithmetic is spelled with explicit roundings so that every kernel that performs it (the
// three-kernel sequence and the fused single-launch update) produces the same bits whatever the compiler contracts.
__device__ __forceinline__ void acc_one(double x, double s, double& a, double& b) {
    const double d = __dsub_rn(x, s);
    a = __dadd_rn(a, d);
    b = __fma_rn(d, d, b);
}
__device__ __forceinline__ double advance_return(double ret, double gamma, double reward) {
    return __fma_rn(ret, gamma, reward);
}
// RunningMeanStd.update_from_moments with the batch given as shifted sums (shift = the running mean itself)
__device__ __forceinline__ void rms_merge_one(double n, double batch, double sa, double sb, double& mean, double& var) {
    const double d1 = __ddiv_rn(sa, batch);  // batch_mean - running mean
    const double d1sq = __dmul_rn(d1, d1);
    const double bvar = fmax(__dsub_rn(__ddiv_rn(sb, batch), d1sq), 0.0);  // population variance of the batch
    const double tot = __dadd_rn(n, batch);
    const double cross = __ddiv_rn(__dmul_rn(__dmul_rn(d1sq, n), batch), tot);
    const double m2 = __dadd_rn(__dadd_rn(__dmul_rn(var, n), __dmul_rn(bvar, batch)), cross);
    mean = __dadd_rn(mean, __ddiv_rn(__dmul_rn(d1, batch), tot));
    var = __ddiv_rn(m2, tot);
}

// partial[(p*kAccBlocks + b)*2 + {0,1}] = sum over this block's strided slice of (x - shift_p), (x - shift_p)^2
__global__ void __launch_bounds__(kAccThreads) acc_partial_kernel(int64_t N, int64_t ld, const double* __restrict__ X,
                                                                  const double* __restrict__ shift,
                                                                  double* __restrict__ partial) {
    const int p = blockIdx.y;
    const double s = shift ? shift[p] : 0.0;
    const double* x = X + (int64_t)p * ld;
    double a = 0.0, b = 0.0;
    // kAccBatch independent loads in flight per thread, accumulated in index order (the summation tree is unchanged)
    constexpr int64_t stride = (int64_t)kAccBlocks * kAccThreads;
    int64_t i = (int64_t)blockIdx.x * kAccThreads + threadIdx.x;
    for (; i + (kAccBatch - 1) * stride < N; i += kAccBatch * stride) {
        double v[kAccBatch];
#pragma unroll
        for (int k = 0; k < kAccBatch; k++) v[k] = x[i + k * stride];
#pragma unroll
        for (int k = 0; k < kAccBatch; k++) acc_one(v[k], s, a, b);
    }
    for (; i < N; i += stride) acc_one(x[i], s, a, b);
    double2 r = block_sum2(a, b);
    if (threadIdx.x == 0) {
        partial[((int64_t)p * kAccBlocks + blockIdx.x) * 2] = r.x;
        partial[((int64_t)p * kAccBlocks + blockIdx.x) * 2 + 1] = r.y;
    }
}
__global__ void acc_final_kernel(int P, const double* __restrict__ partial, double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double a = 0.0, b = 0.0;
    for (int k = 0; k < kAccBlocks; k++) {
        a += partial[((int64_t)p * kAccBlocks + k) * 2];
        b += partial[((int64_t)p * kAccBlocks + k) * 2 + 1];
    }
    out[p] = a;
    out[P + p] = b;
}

// RunningMeanStd.update_from_moments with the batch given as shifted sums (shift = the running mean itself)
__global__ void rms_merge_kernel(int P, double batch_count, const double* __restrict__ sums, double* __restrict__ mean,
                                 double* __restrict__ var, double* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const double n = count[0];
    if (p < P && batch_count > 0) {
        double m = mean[p], v = var[p];
        rms_merge_one(n, batch_count, sums[p], sums[P + p], m, v);
        mean[p] = m;
        var[p] = v;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && batch_count > 0) count[1] = n + batch_count;  // committed by the host-side swap
}
__global__ void rms_commit_kernel(double* __restrict__ count) { count[0] = count[1]; }

// ---- single-rank fast path: accumulate + fold + merge in ONE launch ----------------------------------------------
// Same partial sums as acc_partial_kernel (same slices, same tree), then the block that finishes last folds the
// kAccBlocks partials of every plane in index order and applies the merge - bit-identical to the three-kernel
// sequence accumulate -> merge -> commit, without the three launches (a normalised sdc-v1 step of 2^20 envs is a
// 115 us kernel: five 2-7 us launches per statistic were a fifth of the step).
// RETURNS: the plane is the discounted return, advanced in the same pass: ret <- ret * gamma + reward.
//
// Several ranks (one process per GPU): the same launch also performs the all-reduce of the moment sums, over peer
// memory (NVLink / NVSwitch P2P stores into every rank's exchange region, include/sdcgym.h sdcgym_xchg) instead of an
// NCCL call between an accumulate and a merge kernel: the block that folds this rank's partial sums
//   1. stores its 2P + 1 numbers (shifted sums, shifted square sums, env count) into slot [parity][rank] of EVERY rank's
//      region, fences system-wide and raises flag [parity][rank] = seq there,
//   2. waits until the `world` flags of its own region show seq,
//   3. adds the slots in rank order (the same order on every rank: the normalisers stay bit-identical) and merges.
// Two parities: a rank can be at most one exchange ahead of a peer (it needs the peer's flag of the previous one).
struct XchgDev {
    int world, rank;
    unsigned long long seq;
    int stride;  // doubles per slot
    double* peer[SDCGYM_MAX_RANKS];  // exchange regions of all ranks (peer[rank] = the local one)
};
__device__ __forceinline__ double* xchg_slot(double* region, int world, int stride, int parity, int r) {
    return region + ((size_t)parity * world + r) * stride;
}
__device__ __forceinline__ unsigned long long* xchg_flags(double* region, int world, int stride, int parity) {
    return reinterpret_cast<unsigned long long*>(region + (size_t)2 * world * stride) + (size_t)parity * world;
}

// Wait for a peer's flag with a deadline (~20 s of SM clocks): a rank that never arrives (crashed process) must not
// leave this kernel spinning forever.  Returns false on timeout; the caller then poisons the statistics with NaN so the
// failure is visible in the very next normalised observation.
__device__ __forceinline__ bool xchg_wait(volatile unsigned long long* flag, unsigned long long seq) {
    const long long t0 = clock64();
    while (*flag != seq) {
        if (clock64() - t0 > 40000000000ll) return false;
    }
    return true;
}

template <bool RETURNS, bool DIST>
__global__ void __launch_bounds__(kAccThreads) update_kernel(int P, int64_t N, int64_t ld, const double* __restrict__ X,
                                                             const double* __restrict__ reward, double gamma,
                                                             double* __restrict__ ret, double* mean, double* var,
                                                             double* count, double* __restrict__ partial,
                                                             double* __restrict__ sums, unsigned int* ticket,
                                                             const XchgDev xc) {
    const int p = blockIdx.y;
    const double s = mean[p];
    double a = 0.0, b = 0.0;
    constexpr int64_t stride = (int64_t)kAccBlocks * kAccThreads;
    auto fetch = [&](int64_t i) {
        if (RETURNS) {
            const double x = advance_return(ret[i], gamma, reward[i]);
            ret[i] = x;
            return x;
        }
        return X[(int64_t)p * ld + i];
    };
    int64_t i = (int64_t)blockIdx.x * kAccThreads + threadIdx.x;
    for (; i + (kAccBatch - 1) * stride < N; i += kAccBatch * stride) {
        double v[kAccBatch];
#pragma unroll
        for (int k = 0; k < kAccBatch; k++) v[k] = fetch(i + k * stride);
#pragma unroll
        for (int k = 0; k < kAccBatch; k++) acc_one(v[k], s, a, b);
    }
    for (; i < N; i += stride) acc_one(fetch(i), s, a, b);
    double2 r = block_sum2(a, b);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partial[((int64_t)p * kAccBlocks + blockIdx.x) * 2] = r.x;
        partial[((int64_t)p * kAccBlocks + blockIdx.x) * 2 + 1] = r.y;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const double n = count[0];
    double batch = (double)N;
    __syncthreads();
    if (DIST) {
        // ---- fold the local partials, publish them to every rank, wait for everybody's ----
        const int parity = (int)(xc.seq & 1ull);
        for (int q = threadIdx.x; q < P; q += kAccThreads) {
            double sa = 0.0, sb = 0.0;
            for (int k = 0; k < kAccBlocks; k++) {
                sa += __ldcg(&partial[((int64_t)q * kAccBlocks + k) * 2]);
                sb += __ldcg(&partial[((int64_t)q * kAccBlocks + k) * 2 + 1]);
            }
            for (int r = 0; r < xc.world; r++) {
                double* slot = xchg_slot(xc.peer[r], xc.world, xc.stride, parity, xc.rank);
                slot[q] = sa;
                slot[P + q] = sb;
            }
        }
        if (threadIdx.x == 0)
            for (int r = 0; r < xc.world; r++) xchg_slot(xc.peer[r], xc.world, xc.stride, parity, xc.rank)[2 * P] = batch;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < xc.world) {
            volatile unsigned long long* remote = xchg_flags(xc.peer[threadIdx.x], xc.world, xc.stride, parity) + xc.rank;
            *remote = xc.seq;
        }
        __shared__ int timed_out;
        if (threadIdx.x == 0) timed_out = 0;
        __syncthreads();
        if (threadIdx.x < xc.world) {
            volatile unsigned long long* mine = xchg_flags(xc.peer[xc.rank], xc.world, xc.stride, parity) + threadIdx.x;
            if (!xchg_wait(mine, xc.seq)) timed_out = 1;
        }
        __threadfence_system();
        __syncthreads();
        double* region = xc.peer[xc.rank];
        batch = timed_out ? CUDART_NAN : 0.0;
        for (int r = 0; r < xc.world; r++)
            batch += __ldcv(xchg_slot(region, xc.world, xc.stride, parity, r) + 2 * P);
        for (int q = threadIdx.x; q < P; q += kAccThreads) {
            double sa = 0.0, sb = 0.0;
            for (int r = 0; r < xc.world; r++) {
                const double* slot = xchg_slot(region, xc.world, xc.stride, parity, r);
                sa += __ldcv(slot + q);
                sb += __ldcv(slot + P + q);
            }
            sums[q] = sa;
            sums[P + q] = sb;
            double m = mean[q], v = var[q];
            if (batch > 0.0 || batch != batch) rms_merge_one(n, batch, sa, sb, m, v);  // (NaN batch: a peer timed out)
            mean[q] = m;
            var[q] = v;
        }
    } else {
        for (int q = threadIdx.x; q < P; q += kAccThreads) {
            double sa = 0.0, sb = 0.0;
            for (int k = 0; k < kAccBlocks; k++) {
                sa += __ldcg(&partial[((int64_t)q * kAccBlocks + k) * 2]);
                sb += __ldcg(&partial[((int64_t)q * kAccBlocks + k) * 2 + 1]);
            }
            sums[q] = sa;
            sums[P + q] = sb;
            double m = mean[q], v = var[q];
            rms_merge_one(n, batch, sa, sb, m, v);
            mean[q] = m;
            var[q] = v;
        }
    }
    if (threadIdx.x == 0) {
        count[0] = n + batch;
        count[1] = n + batch;
        *ticket = 0u;  // self-cleaning: ready for the next launch on this stream
    }
}


__global__ void apply_kernel(int64_t N, int64_t ld, const double* __restrict__ X, const double* __restrict__ mean,
                             const double* __restrict__ var, double eps, double clip, double* __restrict__ Y) {
    const int p = blockIdx.y;
    const double m = mean[p], is = 1.0 / sqrt(var[p] + eps);
    // one 8-byte element per thread and iteration: measured 5.6 TB/s (86 % of the copy peak); a 128-bit variant
    // (two envs per access) measured 15 % slower
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = __dmul_rn(__dsub_rn(X[(int64_t)p * ld + i], m), is);
        Y[(int64_t)p * ld + i] = fmin(fmax(v, -clip), clip);
    }
}

__global__ void returns_kernel(int64_t N, const double* __restrict__ reward, double gamma, double* __restrict__ ret) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) ret[i] = advance_return(ret[i], gamma, reward[i]);
}
__global__ void reward_apply_kernel(int64_t N, const double* __restrict__ reward, const uint8_t* __restrict__ flags,
                                    const double* __restrict__ ret_var, double eps, double clip, int normalize,
                                    double* __restrict__ out, double* __restrict__ ret) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double r = reward[i];
    if (normalize) r = fmin(fmax(r / sqrt(ret_var[0] + eps), -clip), clip);
    out[i] = r;
    if (ret && flags && (flags[i] & SDCGYM_FLAG_DONE)) ret[i] = 0.0;
}

// ---- ResidualLoss.take_step (dp_playground.py:247-258): u' = u + P^{-1} r_old, r' = u0 - C u', ||r'||inf -------
template <int M>
struct ResParams {
    double Q[M * M];
    double Qd[M * M];
    int64_t N;
    const double *lam, *qd, *Cs, *u0, *u, *r_old;
    double *u_out, *r_out, *norm_out;
    double* grad;  // optional [N][A][2]: d ||r'||inf / d theta_k as complex g (d = Re(sum g_k d theta_k))
    double dt;
    int32_t prec_type, qd_is_complex, n_act;
};

template <int M>
__global__ void __launch_bounds__(128) residual_step_kernel(const __grid_constant__ ResParams<M> p) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= p.N) return;
    const C2 z{p.lam[2 * i] * p.dt, p.lam[2 * i + 1] * p.dt};
    C2 Qd[M * M];
    const int w = p.qd_is_complex ? 2 : 1;
    const double* row = p.qd ? p.qd + i * (int64_t)p.n_act * w : nullptr;
    int k = 0;
#pragma unroll
    for (int r = 0; r < M; r++)
#pragma unroll
        for (int c = 0; c < M; c++) {
            C2 d{0.0, 0.0};
            bool take = false;
            switch (p.prec_type) {
            case SDCGYM_PREC_DIAG: take = (c == r); break;
            case SDCGYM_PREC_LOWER_DIAG: take = (r == c + 1); break;
            case SDCGYM_PREC_LOWER_TRI: take = (c <= r); break;
            case SDCGYM_PREC_STRICTLY_LOWER_TRI: take = (c < r); break;
            default: break;
            }
            if (p.prec_type == SDCGYM_PREC_FIXED) d.r = (c <= r) ? p.Qd[r * M + c] : 0.0;
            else if (take) {
                d.r = row[k * w];
                if (p.qd_is_complex) d.i = row[k * w + 1];
                k++;
            }
            Qd[r * M + c] = d;
        }
    // delta = P^{-1} r_old by forward substitution, P = I - z Qd (lower triangular)
    C2 un[M], dl[M];
#pragma unroll
    for (int r = 0; r < M; r++) {
        C2 acc{p.r_old[(i * M + r) * 2], p.r_old[(i * M + r) * 2 + 1]};
#pragma unroll
        for (int c = 0; c < r; c++) acc = c_add(acc, c_mul(c_mul(z, Qd[r * M + c]), dl[c]));
        dl[r] = c_div(acc, c_sub(C2{1.0, 0.0}, c_mul(z, Qd[r * M + r])));
        un[r] = c_add(C2{p.u[(i * M + r) * 2], p.u[(i * M + r) * 2 + 1]}, dl[r]);
    }
    double nrm = 0.0;
    bool nan = false;
    int kmax = 0;
    C2 rk{0.0, 0.0};
#pragma unroll
    for (int r = 0; r < M; r++) {
        C2 acc{0.0, 0.0};
#pragma unroll
        for (int c = 0; c < M; c++) {
            C2 cij;
            if (p.Cs) cij = C2{p.Cs[((i * M + r) * M + c) * 2], p.Cs[((i * M + r) * M + c) * 2 + 1]};
            else cij = C2{(r == c ? 1.0 : 0.0) - z.r * p.Q[r * M + c], -z.i * p.Q[r * M + c]};
            acc = c_add(acc, c_mul(cij, un[c]));
        }
        C2 res = c_sub(C2{p.u0[(i * M + r) * 2], p.u0[(i * M + r) * 2 + 1]}, acc);
        p.u_out[(i * M + r) * 2] = un[r].r;
        p.u_out[(i * M + r) * 2 + 1] = un[r].i;
        p.r_out[(i * M + r) * 2] = res.r;
        p.r_out[(i * M + r) * 2 + 1] = res.i;
        double a = hypot(res.r, res.i);
        nan |= isnan(a);
        if (a > nrm) {
            nrm = a;
            kmax = r;
            rk = res;
        }
    }
    p.norm_out[i] = nan ? CUDART_NAN : nrm;
    if (p.grad) {
        // d r' = -z C P^{-1} dQd delta  =>  d ||r'||inf = Re( conj(r'_k)/|r'_k| * (-z) a_i delta_j dQd_ij ),
        // a = P^{-T} C_{k,:}^T (back substitution on the upper-triangular P^T), k = argmax |r'_m|
        C2 a[M];
#pragma unroll
        for (int m = M - 1; m >= 0; m--) {
            C2 ck{0.0, 0.0};
#pragma unroll
            for (int r = 0; r < M; r++)
                if (r == kmax) {
                    if (p.Cs) ck = C2{p.Cs[((i * M + r) * M + m) * 2], p.Cs[((i * M + r) * M + m) * 2 + 1]};
                    else ck = C2{(r == m ? 1.0 : 0.0) - z.r * p.Q[r * M + m], -z.i * p.Q[r * M + m]};
                }
            C2 acc = ck;
#pragma unroll
            for (int q = m + 1; q < M; q++) acc = c_add(acc, c_mul(c_mul(z, Qd[q * M + m]), a[q]));  // - P_qm a_q
            a[m] = c_div(acc, c_sub(C2{1.0, 0.0}, c_mul(z, Qd[m * M + m])));
        }
        C2 f{0.0, 0.0};
        if (nrm > 0.0 && !nan) f = c_mul(c_scale(c_conj(rk), 1.0 / nrm), C2{-z.r, -z.i});
        double* g = p.grad + i * (int64_t)p.n_act * 2;
        int kk = 0;
#pragma unroll
        for (int r = 0; r < M; r++)
#pragma unroll
            for (int c = 0; c < M; c++) {
                bool take = false;
                switch (p.prec_type) {
                case SDCGYM_PREC_DIAG: take = (c == r); break;
                case SDCGYM_PREC_LOWER_DIAG: take = (r == c + 1); break;
                case SDCGYM_PREC_LOWER_TRI: take = (c <= r); break;
                case SDCGYM_PREC_STRICTLY_LOWER_TRI: take = (c < r); break;
                default: break;
                }
                if (take) {
                    const C2 gv = c_mul(f, c_mul(a[r], dl[c]));
                    g[2 * kk] = gv.r;
                    g[2 * kk + 1] = gv.i;
                    kk++;
                }
            }
    }
}

template <int M>
static int launch_residual(const sdcgym_rho_desc* d, int64_t N, const double* lam, const double* qd, const double* Cs,
                           const double* u0, const double* u, const double* r_old, double* u_out, double* r_out,
                           double* norm_out, double* grad, cudaStream_t s) {
    ResParams<M> p;
    for (int k = 0; k < M * M; k++) {
        p.Q[k] = d->Q[k];
        p.Qd[k] = d->Qd_fixed[k];
    }
    p.N = N; p.lam = lam; p.qd = qd; p.Cs = Cs; p.u0 = u0; p.u = u; p.r_old = r_old;
    p.u_out = u_out; p.r_out = r_out; p.norm_out = norm_out; p.grad = grad;
    p.dt = d->dt; p.prec_type = d->prec_type; p.qd_is_complex = d->qd_is_complex;
    p.n_act = sdcgym_num_actions(M, d->prec_type);
    residual_step_kernel<M><<<(unsigned)((N + 127) / 128), 128, 0, s>>>(p);
    return (int)cudaGetLastError();
}

// ---- generalised advantage estimation over a device rollout (SB3 RolloutBuffer.compute_returns_and_advantage) ----
// one thread per env, reverse scan over the T stored steps; every access is coalesced along the env axis.
__global__ void __launch_bounds__(256) gae_kernel(int T, int64_t N, const double* __restrict__ rewards,
                                                  const double* __restrict__ values,
                                                  const uint8_t* __restrict__ episode_starts,
                                                  const double* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, double gamma, double lam,
                                                  double* __restrict__ advantages, double* __restrict__ returns) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    double last_gae = 0.0;
    double next_value = last_values[i];
    double next_non_terminal = last_dones[i] ? 0.0 : 1.0;
    for (int t = T - 1; t >= 0; t--) {
        const int64_t k = (int64_t)t * N + i;
        const double v = values[k];
        const double delta = rewards[k] + gamma * next_value * next_non_terminal - v;
        last_gae = delta + gamma * lam * next_non_terminal * last_gae;
        advantages[k] = last_gae;
        returns[k] = last_gae + v;
        next_value = v;
        next_non_terminal = episode_starts[k] ? 0.0 : 1.0;
    }
}

}  // namespace sdcgym

using namespace sdcgym;

extern "C" int sdcgym_gae(int T, int64_t N, const double* rewards, const double* values, const uint8_t* episode_starts,
                          const double* last_values, const uint8_t* last_dones, double gamma, double gae_lambda,
                          double* advantages, double* returns, void* stream) {
    if (T < 0 || N < 0) return SDCGYM_EINVAL;
    if (T == 0 || N == 0) return 0;
    if (!rewards || !values || !episode_starts || !last_values || !last_dones || !advantages || !returns) return SDCGYM_ENULL;
    gae_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, N, rewards, values, episode_starts,
                                                                            last_values, last_dones, gamma, gae_lambda,
                                                                            advantages, returns);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_scratch_doubles(int P) { return P * kAccBlocks * 2 + 2; }  // partials + the ticket word

// ---- observation planes AND the return plane in one launch (grid.y = P + 1; plane P is the discounted return, advanced
//      in the same pass).  Same slices, same trees, same merges as the two separate launches - bit-identical results -
//      but one launch, one ticket and, with several ranks, ONE exchange for both statistics; the 64 blocks of the
//      return plane no longer have the GPU to themselves. ----
template <bool DIST>
__global__ void __launch_bounds__(kAccThreads) update_both_kernel(int P, int64_t N, int64_t ld, const double* __restrict__ X,
                                                                  const double* __restrict__ reward, double gamma,
                                                                  double* __restrict__ ret, double* omean, double* ovar,
                                                                  double* ocount, double* rmean, double* rvar,
                                                                  double* rcount, double* __restrict__ partial,
                                                                  double* __restrict__ osums, double* __restrict__ rsums,
                                                                  unsigned int* ticket, const XchgDev xc) {
    const int p = blockIdx.y;
    const bool is_ret = (p == P);
    const double s = is_ret ? rmean[0] : omean[p];
    double a = 0.0, b = 0.0;
    constexpr int64_t stride = (int64_t)kAccBlocks * kAccThreads;
    auto fetch = [&](int64_t i) {
        if (is_ret) {
            const double x = advance_return(ret[i], gamma, reward[i]);
            ret[i] = x;
            return x;
        }
        return X[(int64_t)p * ld + i];
    };
    int64_t i = (int64_t)blockIdx.x * kAccThreads + threadIdx.x;
    for (; i + (kAccBatch - 1) * stride < N; i += kAccBatch * stride) {
        double v[kAccBatch];
#pragma unroll
        for (int k = 0; k < kAccBatch; k++) v[k] = fetch(i + k * stride);
#pragma unroll
        for (int k = 0; k < kAccBatch; k++) acc_one(v[k], s, a, b);
    }
    for (; i < N; i += stride) acc_one(fetch(i), s, a, b);
    double2 r = block_sum2(a, b);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partial[((int64_t)p * kAccBlocks + blockIdx.x) * 2] = r.x;
        partial[((int64_t)p * kAccBlocks + blockIdx.x) * 2 + 1] = r.y;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const double on = ocount[0], rn = rcount[0];
    double batch = (double)N;
    const int PP = P + 1;
    __shared__ double fa[4 * SDCGYM_MAX_M + 1], fb[4 * SDCGYM_MAX_M + 1];
    __syncthreads();
    for (int q = threadIdx.x; q < PP; q += kAccThreads) {
        double sa = 0.0, sb = 0.0;
        for (int k = 0; k < kAccBlocks; k++) {
            sa += __ldcg(&partial[((int64_t)q * kAccBlocks + k) * 2]);
            sb += __ldcg(&partial[((int64_t)q * kAccBlocks + k) * 2 + 1]);
        }
        fa[q] = sa;
        fb[q] = sb;
    }
    __syncthreads();
    if (DIST) {
        const int parity = (int)(xc.seq & 1ull);
        for (int q = threadIdx.x; q < PP; q += kAccThreads)
            for (int rk = 0; rk < xc.world; rk++) {
                double* slot = xchg_slot(xc.peer[rk], xc.world, xc.stride, parity, xc.rank);
                slot[q] = fa[q];
                slot[PP + q] = fb[q];
            }
        if (threadIdx.x == 0)
            for (int rk = 0; rk < xc.world; rk++) xchg_slot(xc.peer[rk], xc.world, xc.stride, parity, xc.rank)[2 * PP] = batch;
        __threadfence_system();
        __syncthreads();
        __shared__ int timed_out;
        if (threadIdx.x == 0) timed_out = 0;
        __syncthreads();
        if (threadIdx.x < xc.world) {
            volatile unsigned long long* remote = xchg_flags(xc.peer[threadIdx.x], xc.world, xc.stride, parity) + xc.rank;
            *remote = xc.seq;
            volatile unsigned long long* mine = xchg_flags(xc.peer[xc.rank], xc.world, xc.stride, parity) + threadIdx.x;
            if (!xchg_wait(mine, xc.seq)) timed_out = 1;
        }
        __threadfence_system();
        __syncthreads();
        double* region = xc.peer[xc.rank];
        batch = timed_out ? CUDART_NAN : 0.0;
        for (int rk = 0; rk < xc.world; rk++) batch += __ldcv(xchg_slot(region, xc.world, xc.stride, parity, rk) + 2 * PP);
        for (int q = threadIdx.x; q < PP; q += kAccThreads) {
            double sa = 0.0, sb = 0.0;
            for (int rk = 0; rk < xc.world; rk++) {
                const double* slot = xchg_slot(region, xc.world, xc.stride, parity, rk);
                sa += __ldcv(slot + q);
                sb += __ldcv(slot + PP + q);
            }
            fa[q] = sa;
            fb[q] = sb;
        }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < PP; q += kAccThreads) {
        if (q < P) {
            osums[q] = fa[q];
            osums[P + q] = fb[q];
            double m = omean[q], v = ovar[q];
            if (batch > 0.0 || batch != batch) rms_merge_one(on, batch, fa[q], fb[q], m, v);  // (NaN: a peer timed out)
            omean[q] = m;
            ovar[q] = v;
        } else {
            rsums[0] = fa[q];
            rsums[1] = fb[q];
            double m = rmean[0], v = rvar[0];
            if (batch > 0.0 || batch != batch) rms_merge_one(rn, batch, fa[q], fb[q], m, v);
            rmean[0] = m;
            rvar[0] = v;
        }
    }
    if (threadIdx.x == 0) {
        ocount[0] = on + batch;
        ocount[1] = on + batch;
        rcount[0] = rn + batch;
        rcount[1] = rn + batch;
        *ticket = 0u;
    }
}

extern "C" int sdcgym_vecnorm_update(int P, int64_t N, int64_t ld, const double* X, double* mean, double* var,
                                     double* count2, double* scratch, double* sums, void* stream) {
    if (P < 1 || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!X || !mean || !var || !count2 || !scratch || !sums) return SDCGYM_ENULL;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + (int64_t)P * kAccBlocks * 2);
    update_kernel<false, false><<<dim3(kAccBlocks, P), kAccThreads, 0, (cudaStream_t)stream>>>(
        P, N, ld, X, nullptr, 0.0, nullptr, mean, var, count2, scratch, sums, ticket, XchgDev{});
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_update_returns(int64_t N, const double* reward, double gamma, double* returns, double* mean,
                                             double* var, double* count2, double* scratch, double* sums, void* stream) {
    if (N < 0) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!reward || !returns || !mean || !var || !count2 || !scratch || !sums) return SDCGYM_ENULL;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + (int64_t)kAccBlocks * 2);
    update_kernel<true, false><<<dim3(kAccBlocks, 1), kAccThreads, 0, (cudaStream_t)stream>>>(
        1, N, N, nullptr, reward, gamma, returns, mean, var, count2, scratch, sums, ticket, XchgDev{});
    return (int)cudaGetLastError();
}

// ---- multi-rank: the same single launch with the all-reduce of the moment sums over peer memory ------------------
static int fill_xchg(const sdcgym_xchg* x, int P, XchgDev& d) {
    if (!x) return SDCGYM_ENULL;
    if (x->world < 1 || x->world > SDCGYM_MAX_RANKS || x->rank < 0 || x->rank >= x->world) return SDCGYM_EINVAL;
    if (x->slot_doubles < 2 * P + 1 || x->seq == 0) return SDCGYM_EINVAL;
    d.world = x->world;
    d.rank = x->rank;
    d.seq = x->seq;
    d.stride = x->slot_doubles;
    for (int r = 0; r < x->world; r++) {
        if (!x->peers[r]) return SDCGYM_ENULL;
        d.peer[r] = static_cast<double*>(x->peers[r]);
    }
    return 0;
}

extern "C" size_t sdcgym_xchg_bytes(int world, int slot_doubles) {
    if (world < 1 || slot_doubles < 1) return 0;
    return ((size_t)2 * world * slot_doubles) * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
}

extern "C" int sdcgym_vecnorm_update_dist(int P, int64_t N, int64_t ld, const double* X, double* mean, double* var,
                                          double* count2, double* scratch, double* sums, const sdcgym_xchg* xchg,
                                          void* stream) {
    if (P < 1 || P > kAccThreads * 4 || N < 0 || ld < N) return SDCGYM_EINVAL;
    if ((N > 0 && !X) || !mean || !var || !count2 || !scratch || !sums) return SDCGYM_ENULL;
    XchgDev d{};
    int rc = fill_xchg(xchg, P, d);
    if (rc) return rc;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + (int64_t)P * kAccBlocks * 2);
    // (an empty shard still takes part in the exchange)
    update_kernel<false, true><<<dim3(kAccBlocks, P), kAccThreads, 0, (cudaStream_t)stream>>>(
        P, N, ld, X, nullptr, 0.0, nullptr, mean, var, count2, scratch, sums, ticket, d);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_update_returns_dist(int64_t N, const double* reward, double gamma, double* returns,
                                                  double* mean, double* var, double* count2, double* scratch,
                                                  double* sums, const sdcgym_xchg* xchg, void* stream) {
    if (N < 0) return SDCGYM_EINVAL;
    if ((N > 0 && (!reward || !returns)) || !mean || !var || !count2 || !scratch || !sums) return SDCGYM_ENULL;
    XchgDev d{};
    int rc = fill_xchg(xchg, 1, d);
    if (rc) return rc;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + (int64_t)kAccBlocks * 2);
    update_kernel<true, true><<<dim3(kAccBlocks, 1), kAccThreads, 0, (cudaStream_t)stream>>>(
        1, N, N, nullptr, reward, gamma, returns, mean, var, count2, scratch, sums, ticket, d);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_update_both(int P, int64_t N, int64_t ld, const double* X, const double* reward, double gamma,
                                          double* returns, double* obs_mean, double* obs_var, double* obs_count2,
                                          double* ret_mean, double* ret_var, double* ret_count2, double* scratch,
                                          double* sums_obs, double* sums_ret, const sdcgym_xchg* xchg, void* stream) {
    if (P < 1 || P > 4 * SDCGYM_MAX_M || N < 0 || ld < N) return SDCGYM_EINVAL;
    if ((N > 0 && (!X || !reward || !returns)) || !obs_mean || !obs_var || !obs_count2 || !ret_mean || !ret_var ||
        !ret_count2 || !scratch || !sums_obs || !sums_ret)
        return SDCGYM_ENULL;
    if (!xchg && N == 0) return 0;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + (int64_t)(P + 1) * kAccBlocks * 2);
    const dim3 grid(kAccBlocks, P + 1);
    if (xchg) {
        XchgDev d{};
        int rc = fill_xchg(xchg, P + 1, d);
        if (rc) return rc;
        update_both_kernel<true><<<grid, kAccThreads, 0, (cudaStream_t)stream>>>(
            P, N, ld, X, reward, gamma, returns, obs_mean, obs_var, obs_count2, ret_mean, ret_var, ret_count2, scratch,
            sums_obs, sums_ret, ticket, d);
    } else {
        update_both_kernel<false><<<grid, kAccThreads, 0, (cudaStream_t)stream>>>(
            P, N, ld, X, reward, gamma, returns, obs_mean, obs_var, obs_count2, ret_mean, ret_var, ret_count2, scratch,
            sums_obs, sums_ret, ticket, XchgDev{});
    }
    return (int)cudaGetLastError();
}

// ---- exchange regions: cudaMalloc'd (legacy CUDA IPC needs that), zeroed, shared between the ranks' processes by IPC
//      handle; opening a handle enables peer access to the owning device ----
extern "C" int sdcgym_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) return SDCGYM_ENULL;
    if (bytes == 0) return SDCGYM_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return 0;
}
extern "C" int sdcgym_ipc_open(const unsigned char* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return SDCGYM_ENULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int sdcgym_ipc_close(void* dev_ptr) { return dev_ptr ? (int)cudaIpcCloseMemHandle(dev_ptr) : 0; }
extern "C" int sdcgym_ipc_free(void* dev_ptr) { return dev_ptr ? (int)cudaFree(dev_ptr) : 0; }

extern "C" int sdcgym_vecnorm_accumulate(int P, int64_t N, int64_t ld, const double* X, const double* shift,
                                         double* scratch, double* sums, void* stream) {
    if (P < 1 || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (!X || !scratch || !sums) return SDCGYM_ENULL;
    cudaStream_t s = (cudaStream_t)stream;
    acc_partial_kernel<<<dim3(kAccBlocks, P), kAccThreads, 0, s>>>(N, ld, X, shift, scratch);
    acc_final_kernel<<<(P + 63) / 64, 64, 0, s>>>(P, scratch, sums);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_merge(int P, double batch_count, const double* sums, double* mean, double* var,
                                    double* count2, void* stream) {
    if (P < 1 || P > 1024) return SDCGYM_EINVAL;
    if (!sums || !mean || !var || !count2) return SDCGYM_ENULL;
    cudaStream_t s = (cudaStream_t)stream;
    rms_merge_kernel<<<1, 1024, 0, s>>>(P, batch_count, sums, mean, var, count2);
    if (batch_count > 0) rms_commit_kernel<<<1, 1, 0, s>>>(count2);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_apply(int P, int64_t N, int64_t ld, const double* X, const double* mean, const double* var,
                                    double eps, double clip, double* Y, void* stream) {
    if (P < 1 || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!X || !Y || !mean || !var) return SDCGYM_ENULL;
    unsigned gx = (unsigned)((N + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    apply_kernel<<<dim3(gx, P), 256, 0, (cudaStream_t)stream>>>(N, ld, X, mean, var, eps, clip, Y);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_returns(int64_t N, const double* reward, double gamma, double* returns, void* stream) {
    if (N < 0) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!reward || !returns) return SDCGYM_ENULL;
    returns_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, reward, gamma, returns);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_vecnorm_reward(int64_t N, const double* reward, const uint8_t* flags, const double* ret_var,
                                     double eps, double clip, int normalize, double* out, double* returns, void* stream) {
    if (N < 0) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!reward || !out || (normalize && !ret_var)) return SDCGYM_ENULL;
    reward_apply_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, reward, flags, ret_var, eps, clip,
                                                                                      normalize, out, returns);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_residual_step(const sdcgym_rho_desc* d, int64_t N, const double* lam, const double* qd,
                                    const double* Cs, const double* u0, const double* u, const double* r_old,
                                    double* u_out, double* r_out, double* norm_out, double* grad, void* stream) {
    if (!d) return SDCGYM_ENULL;
    if (!sdcgym_supported(d->M, d->prec_type)) return SDCGYM_EUNSUPPORTED;
    if (N < 0) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!lam || !u0 || !u || !r_old || !u_out || !r_out || !norm_out) return SDCGYM_ENULL;
    if (d->prec_type != SDCGYM_PREC_FIXED && !qd) return SDCGYM_ENULL;
    if (grad && d->prec_type == SDCGYM_PREC_FIXED) return SDCGYM_EUNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    switch (d->M) {
#define C(m) case m: return launch_residual<m>(d, N, lam, qd, Cs, u0, u, r_old, u_out, r_out, norm_out, grad, s);
        C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
#undef C
    }
    return SDCGYM_EUNSUPPORTED;
}
