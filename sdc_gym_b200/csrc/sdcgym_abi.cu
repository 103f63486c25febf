// sdcgym_abi.cu - extern "C" entry points of libsdcgym.so (include/sdcgym.h): argument validation and
// dispatch to the per-M translation units, plus the small layout / reduction / probe kernels.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/sdcgym.h"
#include "exact_math.cuh"

#define SDCGYM_DECL_M(m)                                                                                             \
    extern "C" int sdcgym_launch_reset_m##m(const sdcgym_env_desc*, const sdcgym_state*, const double*, const uint8_t*, \
                                            double*, void*);                                                         \
    extern "C" int sdcgym_launch_step_m##m(const sdcgym_env_desc*, const sdcgym_state*, const sdcgym_step_io*, void*);
SDCGYM_DECL_M(2)
SDCGYM_DECL_M(3)
SDCGYM_DECL_M(4)
SDCGYM_DECL_M(5)
SDCGYM_DECL_M(6)
SDCGYM_DECL_M(7)
SDCGYM_DECL_M(8)
SDCGYM_DECL_M(9)
#undef SDCGYM_DECL_M

namespace sdcgym {

// ---- planes S[4M][ld] <-> reference observation layout obs[N][2][M] complex128 -------------------------
// A block transposes a (4M x 128) tile through shared memory so both sides are coalesced.
constexpr int kTileEnvs = 128;

__global__ void __launch_bounds__(kTileEnvs) export_obs_kernel(int P /*4M*/, int64_t N, int64_t ld,
                                                               const double* __restrict__ S, double* __restrict__ obs) {
    extern __shared__ double tile[];  // [P][kTileEnvs + 1]
    const int64_t e0 = (int64_t)blockIdx.x * kTileEnvs;
    const int n = (int)min((int64_t)kTileEnvs, N - e0);
    for (int p = 0; p < P; p++)
        if ((int)threadIdx.x < n) tile[p * (kTileEnvs + 1) + threadIdx.x] = S[p * ld + e0 + threadIdx.x];
    __syncthreads();
    const int total = n * P;
    double* out = obs + e0 * P;
    for (int k = threadIdx.x; k < total; k += kTileEnvs) {
        int e = k / P, p = k - e * P;
        out[k] = tile[p * (kTileEnvs + 1) + e];
    }
}

__global__ void __launch_bounds__(kTileEnvs) import_obs_kernel(int P, int64_t N, int64_t ld,
                                                               const double* __restrict__ obs, double* __restrict__ S) {
    extern __shared__ double tile[];
    const int64_t e0 = (int64_t)blockIdx.x * kTileEnvs;
    const int n = (int)min((int64_t)kTileEnvs, N - e0);
    const int total = n * P;
    const double* in = obs + e0 * P;
    for (int k = threadIdx.x; k < total; k += kTileEnvs) {
        int e = k / P, p = k - e * P;
        tile[p * (kTileEnvs + 1) + e] = in[k];
    }
    __syncthreads();
    for (int p = 0; p < P; p++)
        if ((int)threadIdx.x < n) S[p * ld + e0 + threadIdx.x] = tile[p * (kTileEnvs + 1) + threadIdx.x];
}

__global__ void refresh_resnorm_kernel(int M, int64_t N, int64_t ld, const double* __restrict__ S,
                                       double* __restrict__ resnorm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double m = -CUDART_INF;
    bool nan = false;
    for (int k = 0; k < M; k++) {
        double a = np_cabs(S[(2 * M + 2 * k) * ld + i], S[(2 * M + 2 * k + 1) * ld + i]);
        nan |= isnan(a);
        m = a > m ? a : m;
    }
    resnorm[i] = nan ? CUDART_NAN : m;
}

// ---- deterministic fp64 sum: fixed 1024-block grid-stride partials, then one block folds them --------
constexpr int kSumBlocks = 1024, kSumThreads = 256;

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sm[kSumThreads / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    v = (threadIdx.x < kSumThreads / 32) ? sm[threadIdx.x] : 0.0;
    if (threadIdx.x < 32)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;
}
__global__ void __launch_bounds__(kSumThreads) sum_partial_kernel(int64_t N, const double* __restrict__ x,
                                                                  double* __restrict__ partials) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kSumThreads + threadIdx.x; i < N; i += (int64_t)kSumBlocks * kSumThreads)
        acc += x[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(kSumThreads) sum_final_kernel(const double* __restrict__ partials,
                                                                double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < kSumBlocks; i += kSumThreads) acc += partials[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[0] = acc;
}

// ---- FP64 pipe peak probe: 8 independent DFMA chains per thread --------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int64_t iters, double* __restrict__ sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int64_t k = 0; k < iters; k++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = __fma_rn(a0, b, c);
            a1 = __fma_rn(a1, b, c);
            a2 = __fma_rn(a2, b, c);
            a3 = __fma_rn(a3, b, c);
            a4 = __fma_rn(a4, b, c);
            a5 = __fma_rn(a5, b, c);
            a6 = __fma_rn(a6, b, c);
            a7 = __fma_rn(a7, b, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}
constexpr int kProbeBlocks = 148 * 8, kProbeThreads = 256;

}  // namespace sdcgym

using namespace sdcgym;

extern "C" int sdcgym_abi_version(void) { return SDCGYM_ABI_VERSION; }

extern "C" int sdcgym_num_actions(int M, int prec_type) {
    if (M < 1) return SDCGYM_EINVAL;
    switch (prec_type) {
    case SDCGYM_PREC_DIAG: return M;
    case SDCGYM_PREC_LOWER_DIAG: return M - 1;
    case SDCGYM_PREC_LOWER_TRI: return M * (M + 1) / 2;
    case SDCGYM_PREC_STRICTLY_LOWER_TRI: return M * (M - 1) / 2;
    case SDCGYM_PREC_FIXED: return 0;
    default: return SDCGYM_EINVAL;
    }
}

extern "C" int sdcgym_supported(int M, int prec_type) {
    return (M >= 2 && M <= SDCGYM_MAX_M && prec_type >= SDCGYM_PREC_DIAG && prec_type <= SDCGYM_PREC_FIXED) ? 1 : 0;
}

static int check_desc_state(const sdcgym_env_desc* d, const sdcgym_state* st) {
    if (!d || !st) return SDCGYM_ENULL;
    if (!sdcgym_supported(d->M, d->prec_type)) return SDCGYM_EUNSUPPORTED;
    if (d->env_kind != SDCGYM_ENV_FULL && d->env_kind != SDCGYM_ENV_STEP) return SDCGYM_EINVAL;
    if (d->blas_variant != SDCGYM_BLAS_SKYLAKEX && d->blas_variant != SDCGYM_BLAS_HASWELL) return SDCGYM_EINVAL;
    if (d->max_iters < 0) return SDCGYM_EINVAL;
    if (st->N < 0 || st->ld < st->N) return SDCGYM_EINVAL;
    if (st->N > 0 && (!st->lam || !st->S || !st->resnorm || !st->niter || !st->episodes || !st->rng_ctr)) return SDCGYM_ENULL;
    return 0;
}

extern "C" int sdcgym_reset(const sdcgym_env_desc* d, const sdcgym_state* st, const double* lam_in, const uint8_t* mask,
                            double* old_states, void* stream) {
    int rc = check_desc_state(d, st);
    if (rc) return rc;
    switch (d->M) {
#define C(m) case m: return sdcgym_launch_reset_m##m(d, st, lam_in, mask, old_states, stream);
        C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
#undef C
    }
    return SDCGYM_EUNSUPPORTED;
}

extern "C" int sdcgym_step(const sdcgym_env_desc* d, const sdcgym_state* st, const sdcgym_step_io* io, void* stream) {
    int rc = check_desc_state(d, st);
    if (rc) return rc;
    if (!io) return SDCGYM_ENULL;
    if (d->reward_strategy < SDCGYM_REW_ITERATION_ONLY || d->reward_strategy > SDCGYM_REW_SMOOTHER_FAST_CONVERGENCE)
        return SDCGYM_EUNSUPPORTED;  // 'spectral_radius' reward: compose sdcgym_spectral_radius on the host side
    if (d->prec_type != SDCGYM_PREC_FIXED && st->N > 0 && !io->action) return SDCGYM_ENULL;
    if (io->old_states && d->autoreset) return SDCGYM_EINVAL;  // collect_states: reset is a separate call
    if (d->sweep_mode != SDCGYM_SWEEP_EXACT && d->sweep_mode != SDCGYM_SWEEP_CERTIFIED) return SDCGYM_EINVAL;
    if (d->sweep_mode == SDCGYM_SWEEP_CERTIFIED && st->N > 0 && (!st->cert || !st->fallback_list || !st->fallback_count))
        return SDCGYM_ENULL;
    if (d->sweep_mode == SDCGYM_SWEEP_CERTIFIED && st->N > INT32_MAX) return SDCGYM_EINVAL;  // int32 fallback list
    switch (d->M) {
#define C(m) case m: return sdcgym_launch_step_m##m(d, st, io, stream);
        C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
#undef C
    }
    return SDCGYM_EUNSUPPORTED;
}

extern "C" int sdcgym_export_obs(int M, int64_t N, int64_t ld, const double* S, double* obs, void* stream) {
    if (M < 1 || M > SDCGYM_MAX_M || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!S || !obs) return SDCGYM_ENULL;
    const int P = 4 * M;
    const unsigned grid = (unsigned)((N + kTileEnvs - 1) / kTileEnvs);
    export_obs_kernel<<<grid, kTileEnvs, P * (kTileEnvs + 1) * sizeof(double), (cudaStream_t)stream>>>(P, N, ld, S, obs);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_export_rows(int P, int64_t N, int64_t ld, const double* X, double* out, void* stream) {
    if (P < 1 || P > 4 * SDCGYM_MAX_M || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!X || !out) return SDCGYM_ENULL;
    const unsigned grid = (unsigned)((N + kTileEnvs - 1) / kTileEnvs);
    export_obs_kernel<<<grid, kTileEnvs, P * (kTileEnvs + 1) * sizeof(double), (cudaStream_t)stream>>>(P, N, ld, X, out);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_import_obs(int M, int64_t N, int64_t ld, const double* obs, double* S, void* stream) {
    if (M < 1 || M > SDCGYM_MAX_M || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!S || !obs) return SDCGYM_ENULL;
    const int P = 4 * M;
    const unsigned grid = (unsigned)((N + kTileEnvs - 1) / kTileEnvs);
    import_obs_kernel<<<grid, kTileEnvs, P * (kTileEnvs + 1) * sizeof(double), (cudaStream_t)stream>>>(P, N, ld, obs, S);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_refresh_resnorm(int M, int64_t N, int64_t ld, const double* S, double* resnorm, void* stream) {
    if (M < 1 || M > SDCGYM_MAX_M || N < 0 || ld < N) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!S || !resnorm) return SDCGYM_ENULL;
    refresh_resnorm_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, N, ld, S, resnorm);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_sum_scratch_doubles(void) { return kSumBlocks; }

extern "C" int sdcgym_sum_f64(int64_t N, const double* x, double* scratch, double* out, void* stream) {
    if (N < 0) return SDCGYM_EINVAL;
    if (!out || !scratch || (N > 0 && !x)) return SDCGYM_ENULL;
    sum_partial_kernel<<<kSumBlocks, kSumThreads, 0, (cudaStream_t)stream>>>(N, x, scratch);
    sum_final_kernel<<<1, kSumThreads, 0, (cudaStream_t)stream>>>(scratch, out);
    return (int)cudaGetLastError();
}

extern "C" int sdcgym_fp64_peak_probe(int64_t iters, double* sink, double* flops_out, void* stream) {
    if (iters <= 0) return SDCGYM_EINVAL;
    if (!sink) return SDCGYM_ENULL;
    fp64_peak_kernel<<<kProbeBlocks, kProbeThreads, 0, (cudaStream_t)stream>>>(iters, sink);
    if (flops_out) *flops_out = 2.0 * 64.0 * (double)iters * (double)kProbeBlocks * (double)kProbeThreads;
    return (int)cudaGetLastError();
}
