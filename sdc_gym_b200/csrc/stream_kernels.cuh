// stream_kernels.cuh - the sdc-v1 single-sweep step (diagonal Q_delta) as a persistent, software-pipelined kernel.
//
// Why.  One sdc-v1 env-step moves 445 B and executes ~600-1000 FP64 instructions (five complex reciprocals, one sweep,
// norms, reward).  The plane layout itself streams at 6.3-6.6 TB/s with one env per thread (tools/plane_stream_bench.cu:
// 29 planes in, 26 out - the copy peak), but step_kernel<.., STEP, ..> reaches 3.9-4.2 TB/s: every warp first waits a
// DRAM round trip for its 29 loads and then works on the FP64 pipe for ~2000 cycles; with 24 warps per SM only ~3 of
// them are in their load phase at any time = 21 KB in flight per SM where 6.5 TB/s needs ~35 KB.
//
// What.  Blocks are persistent and walk over tiles of 256 envs.  The inputs of the NEXT tile (23-24 double planes, 3
// int planes and the 256 action rows: 59-61 KB at M = 5) are fetched by bulk asynchronous copies (cp.async.bulk,
// completion on an mbarrier; SASS UBLKCP) into ONE shared-memory stage while the threads compute the current tile from
// registers: wait -> step_one pulls its env from the stage into registers -> __syncthreads -> the lanes of warp 0 re-arm
// the barrier and issue the next tile's copies (one copy per lane) into the same stage -> compute + store.  The
// arithmetic is step_one<.., STEP, ..> unchanged (StepInputs points it at the stage): results are bit-identical to
// step_kernel.
//
// Measured (B200, 2^20 envs, M = 5): default reward 102 -> 93 us per step (4.2 -> 4.6 TB/s in algorithmic bytes);
// `residual_change` 118 -> 110 us with the norm_init plane (staged like the other inputs) and a single issuing thread,
// -> 104 us with the copies of a stage issued by 28 lanes of warp 0 in parallel.  ncu (profiles/
// ncu_stream_kernels_r02t_summary.json): 1400 warp instructions per 32 envs of which 523 on the FP64 pipe, issue slots
// 44 % busy, FP64 pipe 33 %, `wait` (dependent FP64 chains, 4 warps per scheduler) the top stall: the kernel is bound by
// instruction latency at 16 warps per SM, not by DRAM (4.1 TB/s moved).  128 registers and 2 blocks of 256 threads
// per SM measured best (80-96 registers spill 100-400 bytes and were 13-30 % slower; 128-env tiles equal within 2 %).
// Tried and measured slower (kept as compile-time switches): a private stage + mbarrier per WARP (SDCGYM_STREAM_UNIT=32:
// no block barrier at all, but eight times as many 256-byte copies: 119 vs 105 us) and handing the refill to whichever
// warp pulls its envs out of the stage last instead of a block barrier (SDCGYM_STREAM_HANDOFF=1: 108.5 vs 104.9 us).
#pragma once
#include "bulk_copy.cuh"
#include "step_kernels.cuh"

namespace sdcgym {

#ifdef __CUDACC__
#ifndef SDCGYM_STREAM_TILE
#define SDCGYM_STREAM_TILE 256  // (2 KB per plane and copy; 128-env tiles issue twice as many copies and measured slower)
#endif
constexpr int kStreamTile = SDCGYM_STREAM_TILE;  // envs per tile = threads per block

// A "unit" is the group of threads that shares one stage and one mbarrier: the whole block (kStreamUnit = kStreamTile:
// one __syncthreads per tile between "everybody has pulled its env out of the stage" and "re-arm and refill") or a single
// warp (kStreamUnit = 32: every warp runs its own pipeline over its 32 envs of the tile - 256-byte copies, __syncwarp
// instead of the block barrier, the warps of a block drift apart freely).
#ifndef SDCGYM_STREAM_UNIT
#define SDCGYM_STREAM_UNIT SDCGYM_STREAM_TILE  // (32 measured slower: 119 vs 105 us - eight times as many, 256-byte copies)
#endif
constexpr int kStreamUnit = SDCGYM_STREAM_UNIT;
constexpr int kStreamUnits = kStreamTile / kStreamUnit;
static_assert(kStreamUnit == 32 || kStreamUnit == kStreamTile, "a unit is a warp or the block");
#ifndef SDCGYM_STREAM_ISSUE_LANES
#define SDCGYM_STREAM_ISSUE_LANES 32
#endif
constexpr int kStreamIssueLanes = SDCGYM_STREAM_ISSUE_LANES;  // (1: the single-thread issue, kept for A/B measurements)
#ifndef SDCGYM_STREAM_HANDOFF
#define SDCGYM_STREAM_HANDOFF 0  // (1 measured slower: 108.5 vs 104.9 us - the waiting only moves to the mbarrier)
#endif
constexpr bool kStreamHandoff = SDCGYM_STREAM_HANDOFF != 0 && kStreamIssueLanes == 32;  // (0: __syncthreads + warp 0 issues)

template <int M>
struct StreamStage {
    // byte offsets inside a unit's stage (every bulk copy needs 16-byte aligned addresses and sizes)
    static constexpr int lam = 0;                                   // [2][unit] f64
    static constexpr int S = lam + 2 * kStreamUnit * 8;             // [4M][unit] f64
    static constexpr int resnorm = S + 4 * M * kStreamUnit * 8;     // [unit] f64
    static constexpr int niter = resnorm + kStreamUnit * 8;         // [unit] i32
    static constexpr int episodes = niter + kStreamUnit * 4;        // [unit] i32
    static constexpr int rng_ctr = episodes + kStreamUnit * 4;      // [unit] u32
    static constexpr int norm_init = rng_ctr + kStreamUnit * 4;     // [unit] f64 (residual_change reward only)
    static constexpr int action = norm_init + kStreamUnit * 8;      // [unit][2M] f64 at most (complex actions)
    static constexpr int unit_bytes = action + kStreamUnit * 2 * M * 8;
    static constexpr int bytes = unit_bytes * kStreamUnits;
};

// The first kStreamIssueLanes lanes of a unit: lane 0 arms the barrier with the byte count, then the lanes issue the
// copies between them (copy c by lane c % 32).  One thread issuing all 28 copies of an M = 5 stage serialises them in
// front of its own compute (measured: 110.6 -> 103.7 us per 2^20-env residual_change step).  e0: first env of the unit.
template <int M>
__device__ __forceinline__ void stream_issue_tile(const StepParams<M>& p, unsigned char* stage, unsigned long long* bar,
                                                  int64_t e0, int action_row_bytes, int lane) {
    using L = StreamStage<M>;
    const unsigned plane_b = kStreamUnit * 8, int_b = kStreamUnit * 4;
    const unsigned act_b = (unsigned)(action_row_bytes * kStreamUnit);
    const bool with_ninit = p.strategy == SDCGYM_REW_RESIDUAL_CHANGE && p.norm_init != nullptr;
    if (lane == 0) mbar_expect_tx(bar, (2 + 4 * M + 1 + (with_ninit ? 1 : 0)) * plane_b + 3 * int_b + act_b);
    if (kStreamIssueLanes > 1) __syncwarp();
    constexpr int ncopies = 2 + 4 * M + 6;
    for (int c = lane; c < ncopies; c += kStreamIssueLanes) {
        if (c < 2) bulk_g2s(stage + L::lam + c * plane_b, p.lam + (int64_t)c * p.ld + e0, plane_b, bar);
        else if (c < 2 + 4 * M) bulk_g2s(stage + L::S + (c - 2) * plane_b, p.S + (int64_t)(c - 2) * p.ld + e0, plane_b, bar);
        else if (c == 2 + 4 * M) bulk_g2s(stage + L::resnorm, p.resnorm + e0, plane_b, bar);
        else if (c == 3 + 4 * M) bulk_g2s(stage + L::niter, p.niter + e0, int_b, bar);
        else if (c == 4 + 4 * M) bulk_g2s(stage + L::episodes, p.episodes + e0, int_b, bar);
        else if (c == 5 + 4 * M) bulk_g2s(stage + L::rng_ctr, p.rng_ctr + e0, int_b, bar);
        else if (c == 6 + 4 * M) {
            if (with_ninit) bulk_g2s(stage + L::norm_init, p.norm_init + e0, plane_b, bar);
        } else if (act_b) bulk_g2s(stage + L::action, p.action + e0 * p.a_es, act_b, bar);
    }
}

// Full tiles only: the caller launches step_kernel for the tail envs [tiles * kStreamTile, N).
template <int M, int V, int MINB>
__global__ void __launch_bounds__(kStreamTile, MINB) step_stream_kernel(const __grid_constant__ StepParams<M> p, int64_t tiles,
                                                                        int action_row_bytes) {
    using L = StreamStage<M>;
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ __align__(8) unsigned long long bars[kStreamUnits];
    __shared__ unsigned arrivals;  // (hand-off mode) warps that have pulled their envs out of the stage, all tiles so far
    if (threadIdx.x == 0) arrivals = 0u;
    const int unit = threadIdx.x / kStreamUnit, lane = threadIdx.x % kStreamUnit;
    unsigned char* stage = stages + (size_t)unit * L::unit_bytes;
    unsigned long long* bar = &bars[unit];
    auto unit_sync = [] {
        if (kStreamUnit == 32) __syncwarp();
        else __syncthreads();
    };
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unit_sync();
    int64_t tile = blockIdx.x;
    if (tile >= tiles) return;
    if (lane < kStreamIssueLanes)
        stream_issue_tile<M>(p, stage, bar, tile * kStreamTile + unit * kStreamUnit, action_row_bytes, lane);
    unsigned parity = 0;
    StepInputs in;
    in.lam = reinterpret_cast<const double*>(stage + L::lam);
    in.S = reinterpret_cast<const double*>(stage + L::S);
    in.resnorm = reinterpret_cast<const double*>(stage + L::resnorm);
    in.niter = reinterpret_cast<const int32_t*>(stage + L::niter);
    in.episodes = reinterpret_cast<const int32_t*>(stage + L::episodes);
    in.rng_ctr = reinterpret_cast<const uint32_t*>(stage + L::rng_ctr);
    in.norm_init = reinterpret_cast<const double*>(stage + L::norm_init);  // (read only when the reward needs it)
    in.action = action_row_bytes ? reinterpret_cast<const double*>(stage + L::action) : nullptr;
    in.ld = kStreamUnit;
    in.i = lane;
    for (; tile < tiles; tile += gridDim.x) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        const int64_t next = tile + gridDim.x;
        // step_one pulls every input of its env from the stage into registers in its first basic block and then calls
        // this: the stage is free again, so the next tile's copies fly while this tile is computed and stored
        auto release_and_prefetch = [&]() {
            if (kStreamHandoff && kStreamUnit > 32) {
                // no block barrier: every warp counts itself out of the stage, and the warp that arrives last re-arms
                // and refills it - nobody waits for the slowest warp (the barrier was 14 % of all stall samples)
                __syncwarp();
                unsigned last = 0;
                if ((threadIdx.x & 31) == 0) {
                    __threadfence_block();  // this warp's reads of the stage before its arrival
                    last = ((atomicAdd(&arrivals, 1u) + 1u) % (kStreamTile / 32) == 0u) ? 1u : 0u;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last && next < tiles) {
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads -> async-proxy writes
                    stream_issue_tile<M>(p, stage, bar, next * kStreamTile, action_row_bytes, threadIdx.x & 31);
                }
            } else {
                unit_sync();
                if (lane < kStreamIssueLanes && next < tiles)
                    stream_issue_tile<M>(p, stage, bar, next * kStreamTile + unit * kStreamUnit, action_row_bytes, lane);
            }
        };
        step_one<M, SDCGYM_ENV_STEP, V, false, 0>(p, tile * kStreamTile + threadIdx.x, nullptr, 1, nullptr, 1, &in,
                                                  release_and_prefetch);
    }
}
#endif

}  // namespace sdcgym
