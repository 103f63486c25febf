// step_inst.cu - one translation unit per collocation size: compiled once for every M in 2..9 with
// -DSDCGYM_M=<M> (sdc_gym_b200/build.py) so the eight heavy instantiation sets build in parallel.
// Exposes sdcgym_launch_reset_m<M> / sdcgym_launch_step_m<M>, dispatched from sdcgym_abi.cu.
#include "step_params.cuh"

#ifndef SDCGYM_M
#error "compile with -DSDCGYM_M=<2..9>"
#endif

#define SDCGYM_CAT_(a, b) a##b
#define SDCGYM_CAT(a, b) SDCGYM_CAT_(a, b)

namespace sdcgym {

constexpr int kM = SDCGYM_M;
constexpr int kHoldDiag = HoldPolicy<kM>::diag;
constexpr int kHoldDense = HoldPolicy<kM>::dense;

template <int KIND, int V, bool DENSE>
static cudaError_t launch_step(const StepParams<kM>& p, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.N + kBlock - 1) / kBlock);
    step_kernel<kM, KIND, V, DENSE, (DENSE ? kHoldDense : kHoldDiag)><<<grid, kBlock, 0, s>>>(p);
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
    const bool dense = d->prec_type != SDCGYM_PREC_DIAG;
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
