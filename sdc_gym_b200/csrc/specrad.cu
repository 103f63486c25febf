// specrad.cu - sdcgym_spectral_radius: batched rho(K) kernel launcher (see specrad.cuh, include/sdcgym.h).
#include "specrad_params.cuh"

namespace sdcgym {

#ifndef SDCGYM_RHO_MINB
#define SDCGYM_RHO_MINB 4  // 128 registers, 16 warps/SM: +25 % over 2 blocks/SM (profiles/README.md)
#endif
template <int M>
__global__ void __launch_bounds__(128, (M <= 5) ? SDCGYM_RHO_MINB : 1) rho_kernel(const __grid_constant__ RhoParams<M> p) {
    rho_one<M>(p, (int64_t)blockIdx.x * 128 + threadIdx.x);
}

template <int M>
__global__ void __launch_bounds__(128, (M <= 5) ? 2 : 1) rho_grad_kernel(const __grid_constant__ RhoParams<M> p,
                                                                        double* __restrict__ grad) {
    rho_grad_one<M>(p, (int64_t)blockIdx.x * 128 + threadIdx.x, grad);
}

template <int M>
static int launch_rho_grad(const sdcgym_rho_desc* d, int64_t N, const double* lam, const double* qd, double* rho,
                           double* grad, cudaStream_t s) {
    RhoParams<M> p;
    fill_rho_params<M>(p, d, N, lam, qd, rho);
    rho_grad_kernel<M><<<(unsigned)((N + 127) / 128), 128, 0, s>>>(p, grad);
    return (int)cudaGetLastError();
}

template <int M>
static int launch_rho(const sdcgym_rho_desc* d, int64_t N, const double* lam, const double* qd, double* rho,
                      cudaStream_t s) {
    RhoParams<M> p;
    fill_rho_params<M>(p, d, N, lam, qd, rho);
    rho_kernel<M><<<(unsigned)((N + 127) / 128), 128, 0, s>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace sdcgym

using namespace sdcgym;

extern "C" int sdcgym_spectral_radius(const sdcgym_rho_desc* d, int64_t N, const double* lam, const double* qd,
                                      double* rho, void* stream) {
    if (!d) return SDCGYM_ENULL;
    if (!sdcgym_supported(d->M, d->prec_type)) return SDCGYM_EUNSUPPORTED;
    if (N < 0) return SDCGYM_EINVAL;
    if (N == 0) return 0;  // an empty shard (dist.shard_range allows them) is a no-op, whatever the other arguments
    if (!lam && (d->grid_re <= 0 || d->grid_im <= 0 || d->grid_first < 0 || d->grid_first + N > d->grid_re * d->grid_im))
        return SDCGYM_EINVAL;
    if (!rho) return SDCGYM_ENULL;
    if (d->prec_type != SDCGYM_PREC_FIXED && !qd) return SDCGYM_ENULL;
    if (d->prec_type == SDCGYM_PREC_FIXED) {
        for (int r = 0; r < d->M; r++)
            for (int c = r + 1; c < d->M; c++)
                if (d->Qd_fixed[r * d->M + c] != 0.0) return SDCGYM_EUNSUPPORTED;  // not lower triangular
    }
    cudaStream_t s = (cudaStream_t)stream;
    switch (d->M) {
#define C(m) case m: return launch_rho<m>(d, N, lam, qd, rho, s);
        C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
#undef C
    }
    return SDCGYM_EUNSUPPORTED;
}

extern "C" int sdcgym_spectral_radius_grad(const sdcgym_rho_desc* d, int64_t N, const double* lam, const double* qd,
                                           double* rho, double* grad, void* stream) {
    if (!d) return SDCGYM_ENULL;
    if (!sdcgym_supported(d->M, d->prec_type)) return SDCGYM_EUNSUPPORTED;
    if (d->prec_type == SDCGYM_PREC_FIXED) return SDCGYM_EUNSUPPORTED;  // nothing to differentiate
    if (N < 0 || d->qd_broadcast) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!lam || !qd || !rho || !grad) return SDCGYM_ENULL;
    cudaStream_t s = (cudaStream_t)stream;
    switch (d->M) {
#define C(m) case m: return launch_rho_grad<m>(d, N, lam, qd, rho, grad, s);
        C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
#undef C
    }
    return SDCGYM_EUNSUPPORTED;
}
