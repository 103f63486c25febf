// Experiment (not part of the product): FP64 pipe characteristics on B200 - throughput of DFMA / DMUL / DADD and
// the parallelism (warps per SM x independent chains per warp) needed to saturate the pipe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64mb tools/fp64_microbench.cu && /tmp/fp64mb
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(int iters, double* sink, double b, double c) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) a[i] = __fma_rn(a[i], b, c);
                if (OP == 1) a[i] = __dmul_rn(a[i], b);
                if (OP == 2) a[i] = __dadd_rn(a[i], c);
                if (OP == 3) {  // mixed like the sweep: mul, add, fma rotating
                    if ((r % 3) == 0) a[i] = __dmul_rn(a[i], b);
                    else if ((r % 3) == 1) a[i] = __dadd_rn(a[i], c);
                    else a[i] = __fma_rn(a[i], b, c);
                }
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 123.456) sink[0] = s;
}

template <int OP, int ILP>
void run(const char* name, int warps_per_sm) {
    int threads = 128, blocks_per_sm = warps_per_sm / 4;
    if (warps_per_sm < 4) { threads = 32 * warps_per_sm; blocks_per_sm = 1; }
    int blocks = 148 * blocks_per_sm;
    int iters = 20000 / ILP;
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, ILP><<<blocks, threads>>>(iters, sink, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<OP, ILP><<<blocks, threads>>>(iters, sink, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)iters * 8 * ILP * blocks * threads;  // thread-instructions
    double per_clk_sm = inst / (ms * 1e-3) / 148 / 1.965e9;
    // cycles per dependent op for one warp: iters*8 dependent steps
    printf("%-5s ILP=%d warps/SM=%2d : %6.2f thread-inst/clk/SM (peak 64)  %.3f ms  cyc/dependent-step=%.1f\n", name, ILP,
           warps_per_sm, per_clk_sm, ms, ms * 1e-3 * 1.965e9 / ((double)iters * 8));
    cudaFree(sink);
}

int main() {
    for (int w : {1, 2, 4, 8, 12, 16, 32, 64}) run<0, 1>("DFMA", w);
    for (int w : {4, 8, 12, 16}) run<0, 2>("DFMA", w);
    for (int w : {4, 8, 12, 16}) run<0, 4>("DFMA", w);
    for (int w : {4, 8, 12, 16}) run<0, 8>("DFMA", w);
    for (int w : {8, 16, 64}) run<1, 8>("DMUL", w);
    for (int w : {8, 16, 64}) run<2, 8>("DADD", w);
    for (int w : {8, 12, 16, 64}) run<3, 8>("MIX", w);
    return 0;
}
