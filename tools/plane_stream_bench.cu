// Experiment (GPU): what does the plane layout itself allow?  A kernel that reads PIN planes and writes POUT planes of
// N doubles (plane stride ld), one env per thread like the sdc-v1 step kernel, with no arithmetic to speak of - the
// bandwidth ceiling of the access pattern, to hold the step kernel's 4.2 TB/s against.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/plane_stream_bench tools/plane_stream_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int PIN, int POUT, int MINB>
__global__ void __launch_bounds__(128, MINB) stream_kernel(const double* __restrict__ in, double* __restrict__ out, long N, long ld) {
    const long i = (long)blockIdx.x * 128 + threadIdx.x;
    if (i >= N) return;
    double v[PIN];
#pragma unroll
    for (int p = 0; p < PIN; p++) v[p] = in[p * ld + i];
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < PIN; p++) s += v[p];
#pragma unroll
    for (int p = 0; p < POUT; p++) out[p * ld + i] = s + v[p % PIN];
}

template <int PIN, int POUT, int MINB>
void run(const char* name, const double* in, double* out, long N, long ld) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((N + 127) / 128);
    for (int k = 0; k < 3; k++) stream_kernel<PIN, POUT, MINB><<<grid, 128>>>(in, out, N, ld);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int k = 0; k < reps; k++) stream_kernel<PIN, POUT, MINB><<<grid, 128>>>(in, out, N, ld);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double bytes = (double)N * 8.0 * (PIN + POUT);
    printf("{\"kernel\": \"%s\", \"planes_in\": %d, \"planes_out\": %d, \"blocks_per_sm\": %d, \"envs\": %ld, \"ms\": %.4f, \"GBps\": %.1f}\n",
           name, PIN, POUT, MINB, N, ms, bytes / ms * 1e-6);
}

int main(int argc, char** argv) {
    const long N = argc > 1 ? atol(argv[1]) : (1L << 22), ld = N;
    double *in, *out;
    cudaMalloc(&in, sizeof(double) * 32 * ld);
    cudaMalloc(&out, sizeof(double) * 32 * ld);
    cudaMemset(in, 0, sizeof(double) * 32 * ld);
    run<1, 1, 8>("copy 1 plane", in, out, N, ld);
    run<8, 8, 8>("8 in / 8 out", in, out, N, ld);
    run<29, 26, 6>("sdc-v1 step pattern (29 in / 26 out, 24 warps/SM)", in, out, N, ld);
    run<29, 26, 4>("sdc-v1 step pattern (16 warps/SM)", in, out, N, ld);
    run<29, 26, 8>("sdc-v1 step pattern (32 warps/SM)", in, out, N, ld);
    run<20, 20, 8>("apply pattern per env (20 in / 20 out)", in, out, N, ld);
    run<21, 1, 8>("statistics pattern per env (21 in)", in, out, N, ld);
    return 0;
}
