// stream_kernels.cuh - the sdc-v1 single-sweep step (diagonal Q_delta) as a persistent, software-pipelined kernel.
//
// Why.  One sdc-v1 env-step moves 445 B and executes ~600-1000 FP64 instructions (five complex reciprocals, one sweep,
// norms, reward).  The plane layout itself streams at 6.3-6.6 TB/s with one env per thread (tools/plane_stream_bench.cu:
// 29 planes in, 26 out - the copy peak), but step_kernel<.., STEP, ..> reaches 3.9-4.2 TB/s: every warp first waits a
// DRAM round trip for its 29 loads and then works on the FP64 pipe for ~2000 cycles; with 24 warps per SM only ~3 of
// them are in their load phase at any time = 21 KB in flight per SM where 6.5 TB/s needs ~35 KB.
//
// What.  Blocks are persistent and walk over tiles of 256 envs.  The inputs of the NEXT tile (23 double planes, 3 int
// planes and the 256 action rows: 59 KB at M = 5) are fetched by bulk asynchronous copies (cp.async.bulk, completion
// on an mbarrier; SASS UBLKCP) into ONE shared-memory stage while the threads compute the current tile from
// registers: wait -> step_one pulls its env from the stage into registers -> __syncthreads -> one thread re-arms the
// barrier and issues the next tile's copies into the same stage -> compute + store.  The arithmetic is
// step_one<.., STEP, ..> unchanged (StepInputs points it at the stage): results are bit-identical to step_kernel.
//
// Measured (B200, 2^20 envs, M = 5): default reward 102 -> 93 us per step (4.2 -> 4.6 TB/s in algorithmic bytes);
// `residual_change` 118 -> 121 us while it re-derived the initial residual in every step, 110 -> 108 us with the
// norm_init plane - that variant is bound by its dependent FP64 chains (ten divisions, three logarithms, a norm per
// env), which prefetching cannot shorten.  128 registers and 2 blocks of 256 threads per SM measured best here (80
// registers spill 200 bytes in this formulation; 128-env tiles double the number of copies and were slower).
#pragma once
#include "bulk_copy.cuh"
#include "step_kernels.cuh"

namespace sdcgym {

#ifdef __CUDACC__
#ifndef SDCGYM_STREAM_TILE
#define SDCGYM_STREAM_TILE 256  // (2 KB per plane and copy; 128-env tiles issue twice as many copies and measured slower)
#endif
constexpr int kStreamTile = SDCGYM_STREAM_TILE;  // envs per tile = threads per block

template <int M>
struct StreamStage {
    // byte offsets inside the stage (every bulk copy needs 16-byte aligned addresses and sizes)
    static constexpr int lam = 0;                                   // [2][128] f64
    static constexpr int S = lam + 2 * kStreamTile * 8;             // [4M][128] f64
    static constexpr int resnorm = S + 4 * M * kStreamTile * 8;     // [128] f64
    static constexpr int niter = resnorm + kStreamTile * 8;         // [128] i32
    static constexpr int episodes = niter + kStreamTile * 4;        // [128] i32
    static constexpr int rng_ctr = episodes + kStreamTile * 4;      // [128] u32
    static constexpr int action = rng_ctr + kStreamTile * 4;        // [128][2M] f64 at most (complex actions)
    static constexpr int bytes = action + kStreamTile * 2 * M * 8;
};

// Warp 0: lane 0 arms the barrier with the tile's byte count, then the lanes issue the tile's copies between them (copy
// c by lane c % 32).  One thread issuing all 27 copies of an M = 5 tile serialises them in front of its own compute.
#ifndef SDCGYM_STREAM_ISSUE_LANES
#define SDCGYM_STREAM_ISSUE_LANES 32
#endif
constexpr int kStreamIssueLanes = SDCGYM_STREAM_ISSUE_LANES;  // (1: the single-thread issue, kept for A/B measurements)
template <int M>
__device__ __forceinline__ void stream_issue_tile(const StepParams<M>& p, unsigned char* stage, unsigned long long* bar,
                                                  int64_t tile, int action_row_bytes) {
    using L = StreamStage<M>;
    const int64_t e0 = tile * kStreamTile;
    const unsigned plane_b = kStreamTile * 8, int_b = kStreamTile * 4;
    const unsigned act_b = (unsigned)(action_row_bytes * kStreamTile);
    const int lane = threadIdx.x;
    if (lane == 0) mbar_expect_tx(bar, (2 + 4 * M + 1) * plane_b + 3 * int_b + act_b);
    if (kStreamIssueLanes > 1) __syncwarp();
    constexpr int ncopies = 2 + 4 * M + 5;
    for (int c = lane; c < ncopies; c += kStreamIssueLanes) {
        if (c < 2) bulk_g2s(stage + L::lam + c * plane_b, p.lam + (int64_t)c * p.ld + e0, plane_b, bar);
        else if (c < 2 + 4 * M) bulk_g2s(stage + L::S + (c - 2) * plane_b, p.S + (int64_t)(c - 2) * p.ld + e0, plane_b, bar);
        else if (c == 2 + 4 * M) bulk_g2s(stage + L::resnorm, p.resnorm + e0, plane_b, bar);
        else if (c == 3 + 4 * M) bulk_g2s(stage + L::niter, p.niter + e0, int_b, bar);
        else if (c == 4 + 4 * M) bulk_g2s(stage + L::episodes, p.episodes + e0, int_b, bar);
        else if (c == 5 + 4 * M) bulk_g2s(stage + L::rng_ctr, p.rng_ctr + e0, int_b, bar);
        else if (act_b) bulk_g2s(stage + L::action, p.action + e0 * p.a_es, act_b, bar);
    }
}

// Full tiles only: the caller launches step_kernel for the tail envs [tiles * 128, N).
template <int M, int V, int MINB>
__global__ void __launch_bounds__(kStreamTile, MINB) step_stream_kernel(const __grid_constant__ StepParams<M> p, int64_t tiles,
                                                                        int action_row_bytes) {
    using L = StreamStage<M>;
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int64_t tile = blockIdx.x;
    if (tile >= tiles) return;
    if (threadIdx.x < kStreamIssueLanes) stream_issue_tile<M>(p, stage, &bar, tile, action_row_bytes);
    unsigned parity = 0;
    StepInputs in;
    in.lam = reinterpret_cast<const double*>(stage + L::lam);
    in.S = reinterpret_cast<const double*>(stage + L::S);
    in.resnorm = reinterpret_cast<const double*>(stage + L::resnorm);
    in.niter = reinterpret_cast<const int32_t*>(stage + L::niter);
    in.episodes = reinterpret_cast<const int32_t*>(stage + L::episodes);
    in.rng_ctr = reinterpret_cast<const uint32_t*>(stage + L::rng_ctr);
    in.action = action_row_bytes ? reinterpret_cast<const double*>(stage + L::action) : nullptr;
    in.ld = kStreamTile;
    in.i = threadIdx.x;
    for (; tile < tiles; tile += gridDim.x) {
        mbar_wait(&bar, parity);
        parity ^= 1u;
        const int64_t next = tile + gridDim.x;
        // step_one pulls every input of its env from the stage into registers in its first basic block and then calls
        // this: the stage is free again, so the next tile's copies fly while this tile is computed and stored
        auto release_and_prefetch = [&]() {
            __syncthreads();
            if (threadIdx.x < kStreamIssueLanes && next < tiles) stream_issue_tile<M>(p, stage, &bar, next, action_row_bytes);
        };
        step_one<M, SDCGYM_ENV_STEP, V, false, 0>(p, tile * kStreamTile + threadIdx.x, nullptr, 1, nullptr, 1, &in,
                                                  release_and_prefetch);
    }
}
#endif

}  // namespace sdcgym
