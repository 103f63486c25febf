// Experiment / validation driver (GPU): ddiv_fast / drcp_fast (csrc/fast_div.cuh) against __ddiv_rn / 1.0/x, bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I sdc_gym_b200/csrc tools/div_check.cu -o /tmp/div_check && /tmp/div_check [log2_pairs]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "fast_div.cuh"

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// random double: random sign and mantissa, exponent uniform in [elo, ehi]; `style` forces special mantissas
__device__ double make(uint64_t r, int elo, int ehi, int style) {
    uint64_t mant = r & 0xfffffffffffffull;
    if (style == 1) mant = 0;                              // power of two
    if (style == 2) mant = 0xfffffffffffffull;             // all ones
    if (style == 3) mant &= 0xfffff00000000ull;            // short mantissa
    if (style == 4) mant = (mant & 0xff) | 0x8000000000000ull;
    if (style == 5) mant = 0xfffffffffffffull - (mant & 0xff);
    uint64_t e = (uint64_t)(elo + (int)((r >> 52) % (uint64_t)(ehi - elo + 1)));
    uint64_t bits = ((r >> 63) << 63) | (e << 52) | mant;
    return __longlong_as_double((long long)bits);
}
__device__ __noinline__ double lib_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __noinline__ double lib_rcp(double b) { return 1.0 / b; }
__global__ void check(uint64_t seed, uint64_t per_thread, unsigned long long* bad_div, unsigned long long* bad_rcp,
                      unsigned long long* flagged) {
    uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nd = 0, nr = 0, nf = 0;
    for (uint64_t k = 0; k < per_thread; k++) {
        uint64_t r1 = mix(seed + id * per_thread * 2 + 2 * k), r2 = mix(seed + id * per_thread * 2 + 2 * k + 1);
        int style_a = (int)((r1 >> 40) % 12), style_b = (int)((r2 >> 40) % 12);
        // exponents: mostly the range the kernels see, sometimes the whole window and slightly beyond
        int wide = (k & 7) == 0;
        double a = make(r1, wide ? 600 : 1003, wide ? 1440 : 1043, style_a > 5 ? 0 : style_a);
        double b = make(r2, wide ? 600 : 1003, wide ? 1440 : 1043, style_b > 5 ? 0 : style_b);
        if ((k & 15) == 3) b = a * (1.0 + (double)((int)(r2 & 7) - 3) * 2.220446049250313e-16);  // quotient next to 1
        bool bad = false;
        double q = sdcgym::ddiv_fast(a, b, bad);
        double qr = lib_div(a, b);
        if (!bad && __double_as_longlong(q) != __double_as_longlong(qr)) nd++;
        bool bad2 = false;
        double y = sdcgym::drcp_fast(b, bad2);
        double yr = lib_rcp(b);
        if (!bad2 && __double_as_longlong(y) != __double_as_longlong(yr)) nr++;
        nf += bad ? 1 : 0;
    }
    atomicAdd(bad_div, nd);
    atomicAdd(bad_rcp, nr);
    atomicAdd(flagged, nf);
}
int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 32;
    unsigned long long *d, h[3] = {0, 0, 0};
    cudaMalloc(&d, 24);
    cudaMemset(d, 0, 24);
    const uint64_t threads = 148ull * 8 * 256, total = 1ull << lg, per = (total + threads - 1) / threads;
    check<<<148 * 8, 256>>>(12345, per, d, d + 1, d + 2);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("{\"pairs\": %llu, \"div_mismatch\": %llu, \"rcp_mismatch\": %llu, \"outside_window\": %llu, \"cuda\": \"%s\"}\n",
           (unsigned long long)(per * threads), h[0], h[1], h[2], cudaGetErrorString(e));
    return (h[0] || h[1] || e != cudaSuccess) ? 1 : 0;
}
