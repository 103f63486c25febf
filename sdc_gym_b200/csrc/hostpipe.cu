// hostpipe.cu - host-buffer entry point: one C call that steps a batch whose actions and results live in HOST
// memory, pipelining  H2D(actions) -> step + observation export kernels -> D2H(results)  over three streams in
// chunks, so PCIe traffic in both directions overlaps the kernels (include/sdcgym.h: sdcgym_pipe_*).
//
// This is the C-ABI form of `DummyVecEnv.step(actions)` with numpy arrays on both sides (reference call sites
// rl_playground.py:82,138; dp_playground.py:807).  Device buffers stay caller-owned; the pipe owns only its
// streams and events.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/sdcgym.h"

struct sdcgym_pipe {
    int device;
    int max_chunks;
    cudaStream_t s_in, s_k, s_out, s_out2;  // s_out: observations, s_out2: the small per-env result arrays
    cudaEvent_t* ev_in;   // [max_chunks] actions of chunk c are on the device
    cudaEvent_t* ev_k;    // [max_chunks] kernels of chunk c are done
    cudaEvent_t ev_start;
    // small batches: all results are packed into one device block and leave in ONE transfer (pipe-owned staging)
    unsigned char* pack_dev;
    unsigned char* pack_host;
};

// Batches whose results fit this many bytes take the packed path (six ~7 us transfers become one).
constexpr size_t kPackBytes = 256 * 1024;

struct PackLayout {
    size_t obs, reward, residual, lam, niter, flags, total;  // byte offsets inside the block
};
static PackLayout pack_layout(int M, int64_t N) {
    PackLayout L;
    size_t o = 0;
    L.obs = o;      o += (size_t)N * 4 * M * sizeof(double);
    L.reward = o;   o += (size_t)N * sizeof(double);
    L.residual = o; o += (size_t)N * sizeof(double);
    L.lam = o;      o += (size_t)N * 2 * sizeof(double);
    L.niter = o;    o += ((size_t)N * sizeof(int32_t) + 7) / 8 * 8;
    L.flags = o;    o += ((size_t)N + 7) / 8 * 8;
    L.total = o;
    return L;
}

// one thread per env: observation planes -> reference layout [env][u|r][m][re|im], plus the per-env result scalars
__global__ void pack_results_kernel(int M, int64_t N, int64_t ld, const double* __restrict__ S,
                                    const double* __restrict__ reward, const double* __restrict__ residual,
                                    const double* __restrict__ lam, const int32_t* __restrict__ niter,
                                    const uint8_t* __restrict__ flags, unsigned char* __restrict__ block, PackLayout L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double* obs = reinterpret_cast<double*>(block + L.obs) + i * 4 * M;
    for (int p = 0; p < 4 * M; p++) obs[p] = S[p * ld + i];
    if (reward) reinterpret_cast<double*>(block + L.reward)[i] = reward[i];
    if (residual) reinterpret_cast<double*>(block + L.residual)[i] = residual[i];
    if (lam) {
        reinterpret_cast<double*>(block + L.lam)[2 * i] = lam[2 * i];
        reinterpret_cast<double*>(block + L.lam)[2 * i + 1] = lam[2 * i + 1];
    }
    if (niter) reinterpret_cast<int32_t*>(block + L.niter)[i] = niter[i];
    if (flags) (block + L.flags)[i] = flags[i];
}

#define PIPE_CHECK(x)                       \
    do {                                    \
        cudaError_t e_ = (x);               \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

extern "C" int sdcgym_pipe_create(int max_chunks, sdcgym_pipe** out) {
    if (!out) return SDCGYM_ENULL;
    if (max_chunks < 1 || max_chunks > 1024) return SDCGYM_EINVAL;
    sdcgym_pipe* p = new (std::nothrow) sdcgym_pipe();
    if (!p) return SDCGYM_ENOMEM;
    p->max_chunks = max_chunks;
    p->ev_in = new (std::nothrow) cudaEvent_t[max_chunks];
    p->ev_k = new (std::nothrow) cudaEvent_t[max_chunks];
    if (!p->ev_in || !p->ev_k) return SDCGYM_ENOMEM;
    PIPE_CHECK(cudaGetDevice(&p->device));
    PIPE_CHECK(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    PIPE_CHECK(cudaStreamCreateWithFlags(&p->s_k, cudaStreamNonBlocking));
    PIPE_CHECK(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    PIPE_CHECK(cudaStreamCreateWithFlags(&p->s_out2, cudaStreamNonBlocking));
    PIPE_CHECK(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    for (int c = 0; c < max_chunks; c++) {
        PIPE_CHECK(cudaEventCreateWithFlags(&p->ev_in[c], cudaEventDisableTiming));
        PIPE_CHECK(cudaEventCreateWithFlags(&p->ev_k[c], cudaEventDisableTiming));
    }
    PIPE_CHECK(cudaMalloc((void**)&p->pack_dev, kPackBytes));
    PIPE_CHECK(cudaHostAlloc((void**)&p->pack_host, kPackBytes, cudaHostAllocDefault));
    *out = p;
    return 0;
}

extern "C" int sdcgym_pipe_destroy(sdcgym_pipe* p) {
    if (!p) return 0;
    cudaStreamSynchronize(p->s_in);
    cudaStreamSynchronize(p->s_k);
    cudaStreamSynchronize(p->s_out);
    cudaStreamSynchronize(p->s_out2);
    for (int c = 0; c < p->max_chunks; c++) {
        cudaEventDestroy(p->ev_in[c]);
        cudaEventDestroy(p->ev_k[c]);
    }
    cudaEventDestroy(p->ev_start);
    cudaStreamDestroy(p->s_in);
    cudaStreamDestroy(p->s_k);
    cudaStreamDestroy(p->s_out);
    cudaStreamDestroy(p->s_out2);
    cudaFree(p->pack_dev);
    cudaFreeHost(p->pack_host);
    delete[] p->ev_in;
    delete[] p->ev_k;
    delete p;
    return 0;
}

extern "C" int sdcgym_host_alloc(size_t bytes, void** out) {
    if (!out) return SDCGYM_ENULL;
    return (int)cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
}
extern "C" int sdcgym_host_free(void* p) { return p ? (int)cudaFreeHost(p) : 0; }

static int64_t chunk_begin(int64_t N, int c, int chunks) {
    if (c <= 0) return 0;
    if (c >= chunks) return N;
    const double f = (double)c / (double)chunks;
    return (int64_t)((double)N * f * sqrt(f)) / 32 * 32;
}

// device-side VecNormalize between the step and the download (single stream, in order)
static int run_vecnorm(const sdcgym_env_desc* desc, const sdcgym_state* st, const sdcgym_step_io* dev,
                       const sdcgym_vecnorm* vn, cudaStream_t cs) {
    const int P = 4 * desc->M;
    const int64_t N = st->N;
    int rc = 0;
    if (vn->norm_obs) {
        if (vn->training) {
            rc = sdcgym_vecnorm_update(P, N, st->ld, st->S, vn->obs_mean, vn->obs_var, vn->obs_count2, vn->scratch_obs,
                                       vn->sums_obs, cs);
            if (rc) return rc;
        }
        rc = sdcgym_vecnorm_apply(P, N, st->ld, st->S, vn->obs_mean, vn->obs_var, vn->epsilon, vn->clip_obs,
                                  vn->out_planes, cs);
        if (rc) return rc;
    }
    if (vn->training) {
        rc = sdcgym_vecnorm_update_returns(N, dev->reward, vn->gamma, vn->returns, vn->ret_mean, vn->ret_var,
                                           vn->ret_count2, vn->scratch_ret, vn->sums_ret, cs);
        if (rc) return rc;
    }
    return sdcgym_vecnorm_reward(N, dev->reward, dev->flags, vn->ret_var, vn->epsilon, vn->clip_reward, vn->norm_reward ? 1 : 0,
                                 vn->out_reward, vn->returns, cs);
}

static int pipe_step_impl(sdcgym_pipe* p, const sdcgym_env_desc* desc, const sdcgym_state* st, const sdcgym_step_io* dev,
                          double* obs_dev, const sdcgym_host_io* host, int chunks, const sdcgym_vecnorm* vn,
                          void* caller_stream);

extern "C" int sdcgym_pipe_step(sdcgym_pipe* p, const sdcgym_env_desc* desc, const sdcgym_state* st,
                                const sdcgym_step_io* dev, double* obs_dev, const sdcgym_host_io* host, int chunks,
                                void* caller_stream) {
    return pipe_step_impl(p, desc, st, dev, obs_dev, host, chunks, nullptr, caller_stream);
}

extern "C" int sdcgym_pipe_step_vecnorm(sdcgym_pipe* p, const sdcgym_env_desc* desc, const sdcgym_state* st,
                                        const sdcgym_step_io* dev, double* obs_dev, const sdcgym_host_io* host,
                                        const sdcgym_vecnorm* vn, void* caller_stream) {
    if (!vn) return SDCGYM_ENULL;
    if (!dev || !dev->reward || !dev->flags || !vn->out_reward || !vn->returns || !vn->ret_var) return SDCGYM_ENULL;
    if (vn->norm_obs && (!vn->out_planes || !vn->obs_mean || !vn->obs_var)) return SDCGYM_ENULL;
    if (vn->training && (!vn->ret_mean || !vn->ret_count2 || !vn->scratch_ret || !vn->sums_ret)) return SDCGYM_ENULL;
    if (vn->training && vn->norm_obs && (!vn->obs_count2 || !vn->scratch_obs || !vn->sums_obs)) return SDCGYM_ENULL;
    return pipe_step_impl(p, desc, st, dev, obs_dev, host, 1, vn, caller_stream);
}

static int pipe_step_impl(sdcgym_pipe* p, const sdcgym_env_desc* desc, const sdcgym_state* st, const sdcgym_step_io* dev,
                          double* obs_dev, const sdcgym_host_io* host, int chunks, const sdcgym_vecnorm* vn,
                          void* caller_stream) {
    if (!p || !desc || !st || !dev || !host) return SDCGYM_ENULL;
    const int64_t N = st->N;
    if (N < 0 || chunks < 1) return SDCGYM_EINVAL;
    int cur_dev = -1;
    PIPE_CHECK(cudaGetDevice(&cur_dev));
    if (cur_dev != p->device) {  // streams, events and staging live on the device the pipe was created on
        PIPE_CHECK(cudaSetDevice(p->device));
        const int rc = pipe_step_impl(p, desc, st, dev, obs_dev, host, chunks, vn, caller_stream);
        cudaSetDevice(cur_dev);
        return rc;
    }
    if (chunks > p->max_chunks) chunks = p->max_chunks;
    if (N == 0) return 0;
    const int M = desc->M;
    const int A = sdcgym_num_actions(M, desc->prec_type);
    const int aw = A * (desc->action_is_complex ? 2 : 1);  // doubles per env in the action arrays
    if (A > 0 && (!host->action || !dev->action)) return SDCGYM_ENULL;
    if (host->obs && !obs_dev) return SDCGYM_ENULL;
    cudaStream_t cs = (cudaStream_t)caller_stream;

    // ---- small batch: upload, step, pack, ONE download, scatter on the host.  Runs on the caller's stream. ----
    const PackLayout L = pack_layout(M, N);
    if (L.total <= kPackBytes) {
        double* act_dev = const_cast<double*>(dev->action);
        sdcgym_step_io io = *dev;
        if (A > 0) {
            PIPE_CHECK(cudaMemcpyAsync(act_dev, host->action, sizeof(double) * N * aw, cudaMemcpyHostToDevice, cs));
            io.action_env_stride = aw;
            io.action_comp_stride = desc->action_is_complex ? 2 : 1;
        } else {
            io.action = nullptr;
        }
        int rc = sdcgym_step(desc, st, &io, cs);
        if (rc) return rc;
        const double* obs_src = st->S;
        const double* rew_src = dev->reward;
        if (vn) {
            rc = run_vecnorm(desc, st, dev, vn, cs);
            if (rc) return rc;
            if (vn->norm_obs) obs_src = vn->out_planes;
            rew_src = vn->out_reward;
        }
        pack_results_kernel<<<(unsigned)((N + 127) / 128), 128, 0, cs>>>(
            M, N, st->ld, obs_src, host->reward ? rew_src : nullptr, host->residual ? dev->info_residual : nullptr,
            host->lam ? dev->info_lam : nullptr, host->niter ? dev->info_niter : nullptr,
            host->flags ? dev->flags : nullptr, p->pack_dev, L);
        PIPE_CHECK(cudaGetLastError());
        const size_t first = host->obs ? 0 : L.reward;  // skip the observation part when nobody wants it
        PIPE_CHECK(cudaMemcpyAsync(p->pack_host + first, p->pack_dev + first, L.total - first, cudaMemcpyDeviceToHost, cs));
        PIPE_CHECK(cudaStreamSynchronize(cs));
        const unsigned char* b = p->pack_host;
        if (host->obs) memcpy(host->obs, b + L.obs, (size_t)N * 4 * M * sizeof(double));
        if (host->reward && dev->reward) memcpy(host->reward, b + L.reward, (size_t)N * sizeof(double));
        if (host->residual && dev->info_residual) memcpy(host->residual, b + L.residual, (size_t)N * sizeof(double));
        if (host->lam && dev->info_lam) memcpy(host->lam, b + L.lam, (size_t)N * 2 * sizeof(double));
        if (host->niter && dev->info_niter) memcpy(host->niter, b + L.niter, (size_t)N * sizeof(int32_t));
        if (host->flags && dev->flags) memcpy(host->flags, b + L.flags, (size_t)N);
        return 0;
    }

    if (vn) {
        // normalised large batch: the statistics need the whole batch before any observation can leave, so the
        // chunked overlap does not apply; one pass on the caller's stream (large-batch callers that care about
        // throughput keep observations on the device: VecNormalize.step_tensor / collect_rollouts)
        double* act_dev = const_cast<double*>(dev->action);
        sdcgym_step_io io = *dev;
        if (A > 0) {
            PIPE_CHECK(cudaMemcpyAsync(act_dev, host->action, sizeof(double) * N * aw, cudaMemcpyHostToDevice, cs));
            io.action_env_stride = aw;
            io.action_comp_stride = desc->action_is_complex ? 2 : 1;
        } else {
            io.action = nullptr;
        }
        int rc = sdcgym_step(desc, st, &io, cs);
        if (rc) return rc;
        rc = run_vecnorm(desc, st, dev, vn, cs);
        if (rc) return rc;
        if (host->obs) {
            rc = sdcgym_export_obs(M, N, st->ld, vn->norm_obs ? vn->out_planes : st->S, obs_dev, cs);
            if (rc) return rc;
            PIPE_CHECK(cudaMemcpyAsync(host->obs, obs_dev, (size_t)N * 4 * M * sizeof(double), cudaMemcpyDeviceToHost, cs));
        }
        if (host->reward) PIPE_CHECK(cudaMemcpyAsync(host->reward, vn->out_reward, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, cs));
        if (host->flags) PIPE_CHECK(cudaMemcpyAsync(host->flags, dev->flags, (size_t)N, cudaMemcpyDeviceToHost, cs));
        if (host->niter && dev->info_niter) PIPE_CHECK(cudaMemcpyAsync(host->niter, dev->info_niter, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
        if (host->residual && dev->info_residual) PIPE_CHECK(cudaMemcpyAsync(host->residual, dev->info_residual, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, cs));
        if (host->lam && dev->info_lam) PIPE_CHECK(cudaMemcpyAsync(host->lam, dev->info_lam, (size_t)N * 2 * sizeof(double), cudaMemcpyDeviceToHost, cs));
        PIPE_CHECK(cudaStreamSynchronize(cs));
        return 0;
    }
    // everything the caller enqueued before this call happens before the pipeline starts
    PIPE_CHECK(cudaEventRecord(p->ev_start, cs));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_k, p->ev_start, 0));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_out, p->ev_start, 0));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_out2, p->ev_start, 0));
    // The observations are 80 % of the result bytes and go out chunk by chunk; the five small arrays (37 B/env) are
    // copied once per group of chunks: fewer, larger DMA transfers (each copy costs a few microseconds of set-up).
    const int group = chunks >= 4 ? chunks / 2 : 1;
    int64_t small_lo = 0;

    double* act_dev = const_cast<double*>(dev->action);
    for (int c = 0; c < chunks; c++) {
        // chunk boundaries grow like c^1.5: a small first chunk starts the D2H stream (the bottleneck) early, the
        // later, larger chunks keep the number of transfers low
        const int64_t lo = chunk_begin(N, c, chunks), hi = chunk_begin(N, c + 1, chunks);
        if (hi <= lo) continue;
        const int64_t n = hi - lo;
        if (A > 0) {
            PIPE_CHECK(cudaMemcpyAsync(act_dev + lo * aw, host->action + lo * aw, sizeof(double) * n * aw,
                                       cudaMemcpyHostToDevice, p->s_in));
            PIPE_CHECK(cudaEventRecord(p->ev_in[c], p->s_in));
            PIPE_CHECK(cudaStreamWaitEvent(p->s_k, p->ev_in[c], 0));
        }
        // sub-batch views: planes are addressed base + lo, per-env arrays base + lo (* width)
        sdcgym_state s2 = *st;
        s2.N = n;
        s2.lam += lo; s2.S += lo; s2.resnorm += lo; s2.niter += lo; s2.episodes += lo; s2.rng_ctr += lo;
        if (s2.norm_init) s2.norm_init += lo;
        if (s2.cert) s2.cert += lo;                    // certified sweep mode: per-chunk certificate planes and fallback list
        if (s2.fallback_list) s2.fallback_list += lo;  // (indices are local to the chunk)
        if (s2.phase_list) s2.phase_list += 2 * lo;    // phased dense solve: [2][n] lists of the chunk, inverse planes
        if (s2.phase_pinv) s2.phase_pinv += lo;
        sdcgym_step_io io = *dev;
        io.action = A > 0 ? act_dev + lo * aw : nullptr;
        io.action_env_stride = aw;
        io.action_comp_stride = desc->action_is_complex ? 2 : 1;
        if (io.reward) io.reward += lo;
        if (io.flags) io.flags += lo;
        if (io.info_residual) io.info_residual += lo;
        if (io.info_niter) io.info_niter += lo;
        if (io.info_lam) io.info_lam += 2 * lo;
        if (io.terminal_obs) io.terminal_obs += lo;
        sdcgym_env_desc d2 = *desc;
        d2.env_offset += lo;
        int rc = sdcgym_step(&d2, &s2, &io, p->s_k);
        if (rc) return rc;
        if (host->obs) {
            rc = sdcgym_export_obs(M, n, st->ld, st->S + lo, obs_dev + lo * 4 * M, p->s_k);
            if (rc) return rc;
        }
        PIPE_CHECK(cudaEventRecord(p->ev_k[c], p->s_k));
        PIPE_CHECK(cudaStreamWaitEvent(p->s_out, p->ev_k[c], 0));
#define D2H(strm, from, count, hostp, devp, bytes_per_env)                                                           \
    if ((hostp) && (devp))                                                                                          \
        PIPE_CHECK(cudaMemcpyAsync((char*)(hostp) + (size_t)(from) * (bytes_per_env),                                \
                                   (const char*)(devp) + (size_t)(from) * (bytes_per_env),                          \
                                   (size_t)(count) * (bytes_per_env), cudaMemcpyDeviceToHost, strm));
        D2H(p->s_out, lo, n, host->obs, obs_dev, 4 * M * sizeof(double))
        if ((c + 1) % group == 0 || c + 1 == chunks) {
            PIPE_CHECK(cudaStreamWaitEvent(p->s_out2, p->ev_k[c], 0));
            const int64_t sn = hi - small_lo;
            D2H(p->s_out2, small_lo, sn, host->reward, dev->reward, sizeof(double))
            D2H(p->s_out2, small_lo, sn, host->flags, dev->flags, 1)
            D2H(p->s_out2, small_lo, sn, host->niter, dev->info_niter, sizeof(int32_t))
            D2H(p->s_out2, small_lo, sn, host->residual, dev->info_residual, sizeof(double))
            D2H(p->s_out2, small_lo, sn, host->lam, dev->info_lam, 2 * sizeof(double))
            small_lo = hi;
        }
#undef D2H
    }
    PIPE_CHECK(cudaStreamSynchronize(p->s_out));
    PIPE_CHECK(cudaStreamSynchronize(p->s_out2));
    // later work on the caller's stream must see the new state
    PIPE_CHECK(cudaEventRecord(p->ev_start, p->s_k));
    PIPE_CHECK(cudaStreamWaitEvent(cs, p->ev_start, 0));
    return 0;
}


// =====================================================================================================================
// Result blocks (include/sdcgym.h: sdcgym_block_*): the device-side results of a step and the host arrays the caller
// reads are two copies of ONE contiguous layout, so a step's results leave in a single transfer (small / mid-size
// batches) or in a few large chunked ones (large batches), and land where the host-side arrays already live.
// =====================================================================================================================
static inline uint64_t align256(uint64_t x) { return (x + 255u) / 256u * 256u; }

extern "C" int sdcgym_block_layout_init(int M, int64_t N, sdcgym_block_layout* L) {
    if (!L) return SDCGYM_ENULL;
    if (M < 1 || M > SDCGYM_MAX_M || N < 0) return SDCGYM_EINVAL;
    const uint64_t n = (uint64_t)N;
    uint64_t o = 0;
    L->N = N;
    L->M = M;
    L->reserved = 0;
    L->obs_u = o;    o += align256(n * 2 * M * sizeof(double));
    L->reward = o;   o += align256(n * sizeof(double));
    L->residual = o; o += align256(n * sizeof(double));
    L->lam = o;      o += align256(n * 2 * sizeof(double));
    L->niter = o;    o += align256(n * sizeof(int32_t));
    L->flags = o;    o += align256(n);
    L->obs_r = o;    o += align256(n * 2 * M * sizeof(double));
    L->total = o;
    return 0;
}

namespace {
// SDCGYM_PIPE_TRACE=1: time stamps (CUDA events) of every stage of one chunked step, printed to stderr - the tool behind
// the chunk schedule (tools/e2e_sweep.py); off by default, costs nothing then
struct PipeTrace {
    bool on = false;
    cudaEvent_t start{}, h2d[64], k[64], d2h[64], small[64];
    PipeTrace() {
        const char* e = getenv("SDCGYM_PIPE_TRACE");
        on = e && e[0] == '1';
    }
    void init() {
        if (!on || start) return;
        cudaEventCreate(&start);
        for (int c = 0; c < 64; c++) {
            cudaEventCreate(&h2d[c]);
            cudaEventCreate(&k[c]);
            cudaEventCreate(&d2h[c]);
            cudaEventCreate(&small[c]);
        }
    }
};
PipeTrace g_trace;

// the pipe's streams, events and staging belong to the device it was created on: make that device current for the call
struct DeviceScope {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceScope(int want) {
        int cur = -1;
        err = cudaGetDevice(&cur);
        if (err == cudaSuccess && cur != want) {
            err = cudaSetDevice(want);
            if (err == cudaSuccess) prev = cur;
        }
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Wait for a stream by polling: cudaStreamSynchronize may block on an interrupt (tens of microseconds of wake-up
// latency per call); a step is a few milliseconds at most, so the calling thread spins instead.
inline cudaError_t spin_sync(cudaStream_t s) {
    for (;;) {
        const cudaError_t e = cudaStreamQuery(s);
        if (e != cudaErrorNotReady) return e;
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
}

inline double host_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int auto_chunks(int64_t N) {
    // one transfer up to 32 k envs (<= ~4 MB of results: the copy is shorter than the extra launches and events of a
    // pipeline); beyond that overlap H2D | kernels | D2H in 2..8 growing chunks (profiles/e2e_sweep_r02.jsonl)
    if (N < 32768) return 1;
    if (N < 98304) return 2;
    if (N < 393216) return 4;
    return 8;
}
}  // namespace

extern "C" int sdcgym_pipe_step_block(sdcgym_pipe* p, const sdcgym_env_desc* desc, const sdcgym_state* st,
                                      const sdcgym_block_layout* L, const sdcgym_block_io* bio, const sdcgym_vecnorm* vn,
                                      void* caller_stream) {
    if (!p || !desc || !st || !L || !bio) return SDCGYM_ENULL;
    const int64_t N = st->N;
    const int M = desc->M;
    if (N < 0 || L->N != N || L->M != M) return SDCGYM_EINVAL;
    if (N == 0) return 0;
    if (!bio->dev_block || !bio->host_block) return SDCGYM_ENULL;
    const int A = sdcgym_num_actions(M, desc->prec_type);
    const int aw = A * (desc->action_is_complex ? 2 : 1);  // doubles per env in the action arrays
    if (A > 0 && (!bio->action_host || !bio->action_dev)) return SDCGYM_ENULL;
    if (vn) {
        if (!vn->out_reward || !vn->returns || !vn->ret_var) return SDCGYM_ENULL;
        if (vn->norm_obs && (!vn->out_planes || !vn->obs_mean || !vn->obs_var)) return SDCGYM_ENULL;
        if (vn->training && (!vn->ret_mean || !vn->ret_count2 || !vn->scratch_ret || !vn->sums_ret)) return SDCGYM_ENULL;
        if (vn->training && vn->norm_obs && (!vn->obs_count2 || !vn->scratch_obs || !vn->sums_obs)) return SDCGYM_ENULL;
    }
    DeviceScope scope(p->device);
    if (scope.err != cudaSuccess) return (int)scope.err;
    cudaStream_t cs = (cudaStream_t)caller_stream;
    unsigned char* db = bio->dev_block;
    unsigned char* hb = bio->host_block;
    const bool skip_u = (bio->skip_u & 1) != 0 && !(vn && vn->norm_obs);  // normalised u rows are not constant
    const bool lazy_info = (bio->skip_u & 2) != 0;  // chunked path only: niter / residual / lam stay in the device block
    const int P2 = 2 * M;                                            // planes per observation row (u or r)
    const size_t row_bytes = (size_t)P2 * sizeof(double);

    sdcgym_step_io io;
    io.action = A > 0 ? bio->action_dev : nullptr;
    io.action_env_stride = aw;
    io.action_comp_stride = desc->action_is_complex ? 2 : 1;
    io.reward = reinterpret_cast<double*>(db + L->reward);
    io.flags = db + L->flags;
    io.info_residual = reinterpret_cast<double*>(db + L->residual);
    io.info_niter = reinterpret_cast<int32_t*>(db + L->niter);
    io.info_lam = reinterpret_cast<double*>(db + L->lam);
    io.terminal_obs = bio->terminal_obs;
    io.old_states = nullptr;
    double* obs_u_dev = reinterpret_cast<double*>(db + L->obs_u);
    double* obs_r_dev = reinterpret_cast<double*>(db + L->obs_r);

    int chunks = bio->chunks > 0 ? bio->chunks : auto_chunks(N);
    if (chunks > p->max_chunks) chunks = p->max_chunks;
    if (vn) chunks = 1;  // the statistics need the whole batch before any observation can leave

    if (chunks == 1) {
        // ---- upload, step, export, ONE download; everything in order on the caller's stream ----
        if (A > 0)
            PIPE_CHECK(cudaMemcpyAsync(bio->action_dev, bio->action_host, sizeof(double) * N * aw, cudaMemcpyHostToDevice, cs));
        int rc = sdcgym_step(desc, st, &io, cs);
        if (rc) return rc;
        const double* planes = st->S;
        if (vn) {
            rc = run_vecnorm(desc, st, &io, vn, cs);
            if (rc) return rc;
            if (vn->norm_obs) planes = vn->out_planes;
        }
        if (!skip_u) {
            rc = sdcgym_export_rows(P2, N, st->ld, planes, obs_u_dev, cs);
            if (rc) return rc;
        }
        rc = sdcgym_export_rows(P2, N, st->ld, planes + (size_t)P2 * st->ld, obs_r_dev, cs);
        if (rc) return rc;
        const uint64_t first = skip_u ? L->reward : 0;
        PIPE_CHECK(cudaMemcpyAsync(hb + first, db + first, L->total - first, cudaMemcpyDeviceToHost, cs));
        if (vn)  // the host sees the normalised reward; the device block keeps the raw one
            PIPE_CHECK(cudaMemcpyAsync(hb + L->reward, vn->out_reward, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, cs));
        PIPE_CHECK(spin_sync(cs));
        return 0;
    }

    // ---- large batch: H2D(actions) | step + export kernels | D2H(results), chunk by chunk over four streams ----
    const bool trace = g_trace.on && chunks <= 64;
    bool small_sent[64] = {false};
    const double t_enter = trace ? host_ms() : 0.0;
    if (trace) {
        g_trace.init();
        cudaEventRecord(g_trace.start, cs);
    }
    PIPE_CHECK(cudaEventRecord(p->ev_start, cs));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_k, p->ev_start, 0));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_out, p->ev_start, 0));
    PIPE_CHECK(cudaStreamWaitEvent(p->s_out2, p->ev_start, 0));
    const int group = chunks >= 4 ? chunks / 2 : chunks;  // the five small arrays leave once per group of chunks
    int64_t small_lo = 0;
    for (int c = 0; c < chunks; c++) {
        const int64_t lo = chunk_begin(N, c, chunks), hi = chunk_begin(N, c + 1, chunks);
        if (hi <= lo) continue;
        const int64_t n = hi - lo;
        if (A > 0) {
            PIPE_CHECK(cudaMemcpyAsync(bio->action_dev + lo * aw, bio->action_host + lo * aw, sizeof(double) * n * aw,
                                       cudaMemcpyHostToDevice, p->s_in));
            PIPE_CHECK(cudaEventRecord(p->ev_in[c], p->s_in));
            PIPE_CHECK(cudaStreamWaitEvent(p->s_k, p->ev_in[c], 0));
            if (trace) cudaEventRecord(g_trace.h2d[c], p->s_in);
        }
        sdcgym_state s2 = *st;
        s2.N = n;
        s2.lam += lo; s2.S += lo; s2.resnorm += lo; s2.niter += lo; s2.episodes += lo; s2.rng_ctr += lo;
        if (s2.norm_init) s2.norm_init += lo;
        if (s2.cert) s2.cert += lo;                    // certified sweep mode: per-chunk certificate planes and fallback list
        if (s2.fallback_list) s2.fallback_list += lo;  // (indices are local to the chunk)
        if (s2.phase_list) s2.phase_list += 2 * lo;    // phased dense solve: [2][n] lists of the chunk, inverse planes
        if (s2.phase_pinv) s2.phase_pinv += lo;
        sdcgym_step_io io2 = io;
        if (A > 0) io2.action = bio->action_dev + lo * aw;
        io2.reward += lo;
        io2.flags += lo;
        io2.info_residual += lo;
        io2.info_niter += lo;
        io2.info_lam += 2 * lo;
        if (io2.terminal_obs) io2.terminal_obs += lo;
        sdcgym_env_desc d2 = *desc;
        d2.env_offset += lo;
        int rc = sdcgym_step(&d2, &s2, &io2, p->s_k);
        if (rc) return rc;
        if (!skip_u) {
            rc = sdcgym_export_rows(P2, n, st->ld, st->S + lo, obs_u_dev + lo * P2, p->s_k);
            if (rc) return rc;
        }
        rc = sdcgym_export_rows(P2, n, st->ld, st->S + (size_t)P2 * st->ld + lo, obs_r_dev + lo * P2, p->s_k);
        if (rc) return rc;
        PIPE_CHECK(cudaEventRecord(p->ev_k[c], p->s_k));
        PIPE_CHECK(cudaStreamWaitEvent(p->s_out, p->ev_k[c], 0));
        if (trace) cudaEventRecord(g_trace.k[c], p->s_k);
#define SEG(strm, off, from, count, bytes_per_env)                                                         \
    PIPE_CHECK(cudaMemcpyAsync(hb + (off) + (size_t)(from) * (bytes_per_env), db + (off) + (size_t)(from) * (bytes_per_env), \
                               (size_t)(count) * (bytes_per_env), cudaMemcpyDeviceToHost, strm));
        SEG(p->s_out, L->obs_r, lo, n, row_bytes)
        if (!skip_u) SEG(p->s_out, L->obs_u, lo, n, row_bytes)
        if (trace) cudaEventRecord(g_trace.d2h[c], p->s_out);
        if ((c + 1) % group == 0 || c + 1 == chunks) {
            PIPE_CHECK(cudaStreamWaitEvent(p->s_out2, p->ev_k[c], 0));
            const int64_t sn = hi - small_lo;
            SEG(p->s_out2, L->reward, small_lo, sn, sizeof(double))
            SEG(p->s_out2, L->flags, small_lo, sn, 1)
            if (!lazy_info) {
                SEG(p->s_out2, L->niter, small_lo, sn, sizeof(int32_t))
                SEG(p->s_out2, L->residual, small_lo, sn, sizeof(double))
                SEG(p->s_out2, L->lam, small_lo, sn, 2 * sizeof(double))
            }
            small_lo = hi;
            if (trace) {
                cudaEventRecord(g_trace.small[c], p->s_out2);
                small_sent[c] = true;
            }
        }
#undef SEG
    }
    const double t_issued = trace ? host_ms() : 0.0;
    PIPE_CHECK(spin_sync(p->s_out));
    PIPE_CHECK(spin_sync(p->s_out2));
    if (trace) {
        fprintf(stderr, "[sdcgym pipe] N=%lld chunks=%d  host: all work issued after %.3f ms, streams drained after %.3f ms"
                        "  (device time stamps, ms after the call started: actions on device | kernels done | "
                        "observation rows on host | small arrays on host)\n", (long long)N, chunks, t_issued - t_enter,
                host_ms() - t_enter);
        for (int c = 0; c < chunks; c++) {
            const int64_t lo = chunk_begin(N, c, chunks), hi = chunk_begin(N, c + 1, chunks);
            if (hi <= lo) continue;
            float a = -1.f, b = -1.f, d = -1.f, e = -1.f;
            if (A > 0) cudaEventElapsedTime(&a, g_trace.start, g_trace.h2d[c]);
            cudaEventElapsedTime(&b, g_trace.start, g_trace.k[c]);
            cudaEventElapsedTime(&d, g_trace.start, g_trace.d2h[c]);
            if (small_sent[c]) cudaEventElapsedTime(&e, g_trace.start, g_trace.small[c]);
            fprintf(stderr, "  chunk %2d envs %8lld  %7.3f | %7.3f | %7.3f | %7.3f\n", c, (long long)(hi - lo), a, b, d, e);
        }
    }
    PIPE_CHECK(cudaEventRecord(p->ev_start, p->s_k));
    PIPE_CHECK(cudaStreamWaitEvent(cs, p->ev_start, 0));
    return 0;
}
