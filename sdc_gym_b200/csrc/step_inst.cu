// step_inst.cu - one translation unit per collocation size: compiled once for every M in 2..9 with
// -DSDCGYM_M=<M> (sdc_gym_b200/build.py) so the eight heavy instantiation sets build in parallel.
// Exposes sdcgym_launch_reset_m<M> / sdcgym_launch_step_m<M>, dispatched from sdcgym_abi.cu.
#include <cstdlib>

#include "step_params.cuh"
#include "team_kernels.cuh"

#ifndef SDCGYM_M
#error "compile with -DSDCGYM_M=<2..9>"
#endif

#define SDCGYM_CAT_(a, b) a##b
#define SDCGYM_CAT(a, b) SDCGYM_CAT_(a, b)

namespace sdcgym {

constexpr int kM = SDCGYM_M;
constexpr int kHoldDiag = HoldPolicy<kM>::diag;
constexpr int kHoldDense = HoldPolicy<kM>::dense;

#ifdef SDCGYM_TUNE_VARIANTS
// Tuning knob for experiments (not part of the ABI): SDCGYM_TUNE=<n> selects a (min blocks/SM, C residency)
// variant of the M=5 diagonal full-solve kernel.  0 / unset = the shipped default.
static int tune_variant() {
    const char* e = getenv("SDCGYM_TUNE");
    return e ? atoi(e) : 0;
}
#endif

template <int KIND, int V, bool DENSE>
static cudaError_t launch_step(const StepParams<kM>& p, cudaStream_t s) {
#ifdef SDCGYM_TUNE_VARIANTS
    const unsigned grid = (unsigned)((p.N + kBlock - 1) / kBlock);
    if constexpr (kM == 5 && KIND == SDCGYM_ENV_FULL && !DENSE && V == 0) {
        switch (tune_variant()) {
        case 1: step_kernel<kM, KIND, V, DENSE, 2, 2><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 2: step_kernel<kM, KIND, V, DENSE, 2, 3><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 3: step_kernel<kM, KIND, V, DENSE, 2, 4><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 4: step_kernel<kM, KIND, V, DENSE, 1, 3><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 5: step_kernel<kM, KIND, V, DENSE, 0, 3><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 6: step_kernel<kM, KIND, V, DENSE, 0, 4><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 7: step_kernel<kM, KIND, V, DENSE, 0, 5><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 8: step_kernel<kM, KIND, V, DENSE, 0, 6><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
        case 9: step_kernel<kM, KIND, V, DENSE, 1, 4><<<grid, kBlock, 0, s>>>(p); return cudaGetLastError();
#define SDCGYM_TV(n, HOLD, MINB, BLK)                                                                      \
        case n: step_kernel<kM, KIND, V, DENSE, HOLD, MINB, BLK><<<(unsigned)((p.N + BLK - 1) / BLK), BLK, 0, s>>>(p); \
            return cudaGetLastError();
        SDCGYM_TV(10, 1, 6, 64)
        SDCGYM_TV(11, 0, 8, 64)
        SDCGYM_TV(12, 2, 4, 64)
        SDCGYM_TV(13, 1, 12, 32)
        SDCGYM_TV(14, 1, 5, 64)
        SDCGYM_TV(15, 1, 2, 256)
        SDCGYM_TV(16, 0, 2, 256)
        SDCGYM_TV(17, 1, 4, 96)
#define SDCGYM_TVS(n, HOLD, MINB, BLK)                                                                       \
        case n:                                                                                              \
            cudaFuncSetAttribute(step_kernel<kM, KIND, V, DENSE, HOLD, MINB, BLK>,                            \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_kernel_smem_bytes<kM, HOLD, BLK>()); \
            step_kernel<kM, KIND, V, DENSE, HOLD, MINB, BLK><<<(unsigned)((p.N + BLK - 1) / BLK), BLK,        \
                                                               step_kernel_smem_bytes<kM, HOLD, BLK>(), s>>>(p); \
            return cudaGetLastError();
        SDCGYM_TVS(20, 3, 3, 128)
        SDCGYM_TVS(21, 3, 4, 128)
        SDCGYM_TVS(22, 3, 6, 64)
        SDCGYM_TVS(23, 3, 2, 128)
        SDCGYM_TVS(24, 4, 4, 128)
        SDCGYM_TVS(25, 4, 3, 128)
        SDCGYM_TVS(26, 4, 8, 64)
        SDCGYM_TVS(27, 4, 5, 128)
        SDCGYM_TVS(28, 4, 5, 96)
        SDCGYM_TVS(30, 6, 4, 128)
        SDCGYM_TVS(31, 6, 5, 128)
        SDCGYM_TVS(32, 6, 5, 96)
        SDCGYM_TVS(33, 6, 6, 64)
        SDCGYM_TVS(34, 6, 6, 96)
        SDCGYM_TVS(35, 6, 9, 64)
        SDCGYM_TVS(36, 6, 7, 64)
        SDCGYM_TVS(37, 4, 7, 64)
        SDCGYM_TVS(40, 7, 4, 128)
        SDCGYM_TVS(41, 7, 5, 128)
        SDCGYM_TVS(42, 7, 5, 96)
        SDCGYM_TVS(43, 7, 8, 64)
        SDCGYM_TVS(44, 7, 3, 128)
        SDCGYM_TVS(45, 7, 6, 64)
#undef SDCGYM_TVS
#undef SDCGYM_TV
        default: break;
        }
    }
#endif
#ifdef SDCGYM_TUNE_VARIANTS
    if constexpr (kM == 5 && KIND == SDCGYM_ENV_STEP && !DENSE && V == 0) {
        switch (tune_variant()) {
#define SDCGYM_TV1(n, MINB, BLK)                                                                        \
        case n: step_kernel<kM, KIND, V, DENSE, 0, MINB, BLK><<<(unsigned)((p.N + BLK - 1) / BLK), BLK, 0, s>>>(p); \
            return cudaGetLastError();
        SDCGYM_TV1(1, 2, 128)
        SDCGYM_TV1(2, 3, 128)
        SDCGYM_TV1(3, 4, 128)
        SDCGYM_TV1(4, 5, 128)
        SDCGYM_TV1(5, 6, 128)
        SDCGYM_TV1(6, 8, 128)
        SDCGYM_TV1(7, 4, 256)
        SDCGYM_TV1(8, 12, 64)
        SDCGYM_TV1(9, 16, 64)
#undef SDCGYM_TV1
        default: break;
        }
    }
#endif
    if constexpr (DENSE && kM >= kTeamMinM) {
        // large dense Q_delta: one env per team of M lanes (team_kernels.cuh); collect_states keeps the per-thread kernel
        static const bool no_team = getenv("SDCGYM_NO_TEAM") != nullptr;  // A/B switch for tools/bench_dense.py
        if (p.old_states == nullptr && !no_team) {
            constexpr int envs_per_block = (kTeamBlock / 32) * (32 / kM);
#ifdef SDCGYM_TUNE_VARIANTS
            static const int tt = getenv("SDCGYM_TEAM_TUNE") ? atoi(getenv("SDCGYM_TEAM_TUNE")) : 0;
            if (tt == 2) { team_step_kernel<kM, KIND, V, 2><<<(unsigned)((p.N + envs_per_block - 1) / envs_per_block), kTeamBlock, 0, s>>>(p); return cudaGetLastError(); }
            if (tt == 4) { team_step_kernel<kM, KIND, V, 4><<<(unsigned)((p.N + envs_per_block - 1) / envs_per_block), kTeamBlock, 0, s>>>(p); return cudaGetLastError(); }
            if (tt == 5) { team_step_kernel<kM, KIND, V, 5><<<(unsigned)((p.N + envs_per_block - 1) / envs_per_block), kTeamBlock, 0, s>>>(p); return cudaGetLastError(); }
#endif
            team_step_kernel<kM, KIND, V><<<(unsigned)((p.N + envs_per_block - 1) / envs_per_block), kTeamBlock, 0, s>>>(p);
            return cudaGetLastError();
        }
    }
    constexpr bool kStep = (KIND == SDCGYM_ENV_STEP);
    constexpr int hold = DENSE ? kHoldDense : (kStep ? HoldPolicy<kM>::step : kHoldDiag);
    constexpr int minb = DENSE ? HoldPolicy<kM>::dense_minb : (kStep ? HoldPolicy<kM>::step_minb : HoldPolicy<kM>::diag_minb);
    constexpr int block = DENSE ? HoldPolicy<kM>::dense_block : ((!kStep) ? HoldPolicy<kM>::diag_block : kBlock);
    static_assert(!(DENSE && kStep && hold >= 3) || true, "");
    constexpr size_t smem = step_kernel_smem_bytes<kM, hold, block>();
    auto kernel = step_kernel<kM, KIND, V, DENSE, hold, minb, block>;
    if (smem > 48 * 1024) {
        static bool configured = false;  // per instantiation; one host thread per GPU drives the library
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = true;
        }
    }
    kernel<<<(unsigned)((p.N + block - 1) / block), block, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace sdcgym

using namespace sdcgym;

extern "C" int SDCGYM_CAT(sdcgym_launch_reset_m, SDCGYM_M)(const sdcgym_env_desc* d, const sdcgym_state* st,
                                                          const double* lam_in, const uint8_t* mask, double* old_states,
                                                          void* stream) {
    StepParams<kM> p;
    fill_params<kM>(p, d, st);
    p.lam_in = lam_in;
    p.mask = mask;
    p.old_states = old_states;
    if (p.N <= 0) return 0;
    const unsigned grid = (unsigned)((p.N + kBlock - 1) / kBlock);
    cudaStream_t s = (cudaStream_t)stream;
    if (d->blas_variant == SDCGYM_BLAS_SKYLAKEX) reset_kernel<kM, 0><<<grid, kBlock, 0, s>>>(p);
    else reset_kernel<kM, 1><<<grid, kBlock, 0, s>>>(p);
    return (int)cudaGetLastError();
}

extern "C" int SDCGYM_CAT(sdcgym_launch_step_m, SDCGYM_M)(const sdcgym_env_desc* d, const sdcgym_state* st,
                                                         const sdcgym_step_io* io, void* stream) {
    StepParams<kM> p;
    fill_params<kM>(p, d, st);
    fill_step_io<kM>(p, io);
    if (p.N <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    bool dense = d->prec_type != SDCGYM_PREC_DIAG;
    if (d->prec_type == SDCGYM_PREC_FIXED) {  // a fixed diagonal matrix takes the diagonal fast path
        bool diagonal = true;
        for (int r = 0; r < kM; r++)
            for (int c = 0; c < kM; c++)
                if (r != c && d->Qd_fixed[r * kM + c] != 0.0) diagonal = false;
        dense = !diagonal;
    }
    const bool full = d->env_kind == SDCGYM_ENV_FULL;
    const bool skx = d->blas_variant == SDCGYM_BLAS_SKYLAKEX;
    cudaError_t e;
#define SDCGYM_DISPATCH(KIND, V, DENSE) e = launch_step<KIND, V, DENSE>(p, s)
    if (full) {
        if (skx) { if (dense) SDCGYM_DISPATCH(SDCGYM_ENV_FULL, 0, true); else SDCGYM_DISPATCH(SDCGYM_ENV_FULL, 0, false); }
        else     { if (dense) SDCGYM_DISPATCH(SDCGYM_ENV_FULL, 1, true); else SDCGYM_DISPATCH(SDCGYM_ENV_FULL, 1, false); }
    } else {
        if (skx) { if (dense) SDCGYM_DISPATCH(SDCGYM_ENV_STEP, 0, true); else SDCGYM_DISPATCH(SDCGYM_ENV_STEP, 0, false); }
        else     { if (dense) SDCGYM_DISPATCH(SDCGYM_ENV_STEP, 1, true); else SDCGYM_DISPATCH(SDCGYM_ENV_STEP, 1, false); }
    }
#undef SDCGYM_DISPATCH
    return (int)e;
}
