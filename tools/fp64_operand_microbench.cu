// Experiment (not part of the product): does FP64 throughput on B200 depend on how many *register* operands an
// instruction reads?  (The sweep kernel's DFMA/DMUL/DADD read 2-3 distinct 64-bit registers each; the peak probe
// reads one register and two constants.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64ops tools/fp64_operand_microbench.cu && /tmp/fp64ops
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, double* sink, const double* init) {
    constexpr int ILP = 8;
    double a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        a[i] = init[i] + threadIdx.x * 1e-9;
        b[i] = init[8 + i];
        c[i] = init[16 + i];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) a[i] = __fma_rn(a[i], 1.0000001, 1e-9);              // 1 register operand
                if (MODE == 1) a[i] = __fma_rn(a[i], b[i], 1e-9);                   // 2 register operands
                if (MODE == 2) a[i] = __fma_rn(a[i], b[i], c[i]);                   // 3 register operands
                if (MODE == 3) a[i] = __fma_rn(b[i], c[i], a[i]);                   // 3 regs, accumulator form
                if (MODE == 4) a[i] = __dmul_rn(a[i], b[i]);                        // DMUL 2 regs
                if (MODE == 5) a[i] = __dadd_rn(a[i], b[i]);                        // DADD 2 regs
                if (MODE == 6) a[i] = __fma_rn(b[(i + r) & 7], c[(i + 2 * r) & 7], a[i]);  // 3 regs, rotating sources
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 123.456) sink[0] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm, const double* init) {
    int blocks = 148 * (warps_per_sm / 4), iters = 4000;
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 128>>>(iters, sink, init);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128>>>(iters, sink, init);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)iters * 32 * blocks * 128;
    printf("%-28s warps/SM=%2d : %6.2f thread-inst/clk/SM (peak 64)\n", name, warps_per_sm, inst / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(sink);
}

int main() {
    double h[24];
    for (int i = 0; i < 24; i++) h[i] = 1.0 + 1e-7 * i;
    double* init; cudaMalloc(&init, sizeof h); cudaMemcpy(init, h, sizeof h, cudaMemcpyHostToDevice);
    for (int w : {8, 16}) {
        run<0>("DFMA 1 reg + 2 const", w, init);
        run<1>("DFMA 2 reg + 1 const", w, init);
        run<2>("DFMA 3 reg (a=a*b+c)", w, init);
        run<3>("DFMA 3 reg (a=b*c+a)", w, init);
        run<6>("DFMA 3 reg rotating", w, init);
        run<4>("DMUL 2 reg", w, init);
        run<5>("DADD 2 reg", w, init);
    }
    return 0;
}
