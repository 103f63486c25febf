/*
 * tests/c_abi/abi_roundtrip.c - the C ABI used from plain C: no Python, no torch.
 *
 * Builds against include/sdcgym.h + libsdcgym.so + libcudart and against the CPU oracle (oracle/sdc_exact.c, test
 * infrastructure).  Steps a batch of sdc-v0 envs through the host-buffer entry point sdcgym_pipe_step with
 * page-locked host arrays and plain cudaMalloc'ed device buffers, then checks iteration counts, residual norms
 * and terminal states bit for bit against the oracle.  Exit code 0 = parity.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include tests/c_abi/abi_roundtrip.c oracle/sdc_exact.c \
 *       -ffp-contract=off -mfma -L sdc_gym_b200 -lsdcgym -L /usr/local/cuda/lib64 -lcudart -lm -o abi_roundtrip
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdcgym.h"

/* oracle (oracle/sdc_exact.c) */
void sdc_oracle_reset(int M, const double* Q, double dt, int64_t N, const double* lam, double* u, double* r, int variant);
void sdc_oracle_step_v0(int M, const double* Q, double dt, int64_t N, int prec_type, const double* Qd_fixed,
                        const double* action, int action_is_complex, int do_scale, const double* lam, double* u,
                        double* r, int32_t* niter, const double* rinit, int strategy, double step_penalty,
                        double residual_weight, double norm_factor, double restol, int max_iters, double* reward,
                        uint8_t* converged, double* resnorm, uint8_t* err, int variant, double* old_states);

#define CK(x)                                                            \
    do {                                                                 \
        int rc_ = (int)(x);                                              \
        if (rc_ != 0) {                                                  \
            fprintf(stderr, "%s failed: %d (line %d)\n", #x, rc_, __LINE__); \
            return 2;                                                    \
        }                                                                \
    } while (0)

static uint64_t rng_state = 88172645463325252ull;
static double uniform01(void) { /* xorshift64*, deterministic inputs */
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
    const int M = 5;
    const int64_t N = argc > 1 ? atoll(argv[1]) : 20000, ld = (N + 31) / 32 * 32;
    /* Radau IIA(5) collocation matrix is an input: read it from the file the Python test writes */
    double Q[25];
    FILE* f = fopen(argc > 2 ? argv[2] : "Q5.bin", "rb");
    if (!f || fread(Q, sizeof(double), 25, f) != 25) { fprintf(stderr, "cannot read Q\n"); return 2; }
    fclose(f);
    const double xmin[5] = {0.2818591930905709, 0.2011358490453793, 0.06274536689514164, 0.11790265267514095,
                            0.1571629578515223};

    sdcgym_env_desc d;
    memset(&d, 0, sizeof d);
    d.M = M; d.env_kind = SDCGYM_ENV_FULL; d.prec_type = SDCGYM_PREC_DIAG; d.do_scale = 1; d.max_iters = 50;
    d.reward_strategy = SDCGYM_REW_ITERATION_ONLY; d.blas_variant = SDCGYM_BLAS_SKYLAKEX; d.autoreset = 1;
    d.dt = 1.0; d.restol = 1e-10; d.step_penalty = 0.1; d.residual_weight = 0.5; d.norm_factor = 1.0;
    d.lam_re_lo = -100; d.lam_re_hi = 0; d.lam_im_lo = -10; d.lam_im_hi = 0; d.seed = 3;
    memcpy(d.Q, Q, sizeof Q);

    sdcgym_state st;
    memset(&st, 0, sizeof st);
    st.N = N; st.ld = ld;
    CK(cudaMalloc((void**)&st.lam, 2 * ld * 8)); CK(cudaMalloc((void**)&st.S, 4 * M * ld * 8));
    CK(cudaMalloc((void**)&st.resnorm, ld * 8)); CK(cudaMalloc((void**)&st.niter, ld * 4));
    CK(cudaMalloc((void**)&st.episodes, ld * 4)); CK(cudaMalloc((void**)&st.rng_ctr, ld * 4));
    CK(cudaMemset(st.episodes, 0, ld * 4)); CK(cudaMemset(st.rng_ctr, 0, ld * 4));

    sdcgym_step_io dev;
    memset(&dev, 0, sizeof dev);
    double *act_dev, *obs_dev;
    CK(cudaMalloc((void**)&act_dev, N * M * 8)); CK(cudaMalloc((void**)&obs_dev, N * 4 * M * 8));
    dev.action = act_dev; dev.action_env_stride = M; dev.action_comp_stride = 1;
    CK(cudaMalloc((void**)&dev.reward, N * 8)); CK(cudaMalloc((void**)&dev.flags, N));
    CK(cudaMalloc((void**)&dev.info_residual, N * 8)); CK(cudaMalloc((void**)&dev.info_niter, N * 4));
    CK(cudaMalloc((void**)&dev.info_lam, N * 16)); CK(cudaMalloc((void**)&dev.terminal_obs, 4 * M * ld * 8));

    sdcgym_host_io h;
    double *act_h, *term_planes;
    CK(sdcgym_host_alloc(N * M * 8, (void**)&act_h)); CK(sdcgym_host_alloc(N * 4 * M * 8, (void**)&h.obs));
    CK(sdcgym_host_alloc(N * 8, (void**)&h.reward)); CK(sdcgym_host_alloc(N, (void**)&h.flags));
    CK(sdcgym_host_alloc(N * 4, (void**)&h.niter)); CK(sdcgym_host_alloc(N * 8, (void**)&h.residual));
    CK(sdcgym_host_alloc(N * 16, (void**)&h.lam));
    h.action = act_h;
    for (int64_t i = 0; i < N; i++)
        for (int m = 0; m < M; m++) act_h[i * M + m] = 2 * (xmin[m] + (uniform01() - 0.5) * 0.06) - 1;

    /* env.reset(): lambdas from the device Philox stream; fetch them for the oracle */
    CK(sdcgym_reset(&d, &st, NULL, NULL, NULL, NULL));
    double* lam_planes = malloc(2 * ld * 8);
    CK(cudaMemcpy(lam_planes, st.lam, 2 * ld * 8, cudaMemcpyDeviceToHost));
    double* lam = malloc(N * 16);
    for (int64_t i = 0; i < N; i++) { lam[2 * i] = lam_planes[i]; lam[2 * i + 1] = lam_planes[ld + i]; }

    /* env.step(actions) through the host-buffer pipeline */
    sdcgym_pipe* pipe;
    CK(sdcgym_pipe_create(16, &pipe));
    CK(sdcgym_pipe_step(pipe, &d, &st, &dev, obs_dev, &h, 4, NULL));
    term_planes = malloc(4 * M * ld * 8);
    CK(cudaMemcpy(term_planes, dev.terminal_obs, 4 * M * ld * 8, cudaMemcpyDeviceToHost));

    /* oracle */
    double *u = malloc(N * M * 16), *r = malloc(N * M * 16), *rinit = malloc(N * M * 16);
    double *rew = malloc(N * 8), *res = malloc(N * 8);
    int32_t* nit = calloc(N, 4);
    uint8_t *conv = malloc(N), *err = malloc(N);
    sdc_oracle_reset(M, Q, 1.0, N, lam, u, r, 0);
    memcpy(rinit, r, N * M * 16);
    sdc_oracle_step_v0(M, Q, 1.0, N, 0, NULL, act_h, 0, 1, lam, u, r, nit, rinit, 0, 0.1, 0.5, 1.0, 1e-10, 50, rew, conv,
                       res, err, 0, NULL);
    int64_t bad = 0, nconv = 0;
    for (int64_t i = 0; i < N; i++) {
        int ok = nit[i] == h.niter[i] && res[i] == h.residual[i] && rew[i] == h.reward[i] &&
                 lam[2 * i] == h.lam[2 * i] && lam[2 * i + 1] == h.lam[2 * i + 1] &&
                 ((h.flags[i] & SDCGYM_FLAG_CONVERGED) != 0) == (conv[i] != 0) && (h.flags[i] & SDCGYM_FLAG_DONE);
        for (int m = 0; m < M && ok; m++) {
            ok = ok && term_planes[(2 * m) * ld + i] == u[(i * M + m) * 2] && term_planes[(2 * m + 1) * ld + i] == u[(i * M + m) * 2 + 1];
            ok = ok && term_planes[(2 * M + 2 * m) * ld + i] == r[(i * M + m) * 2] &&
                 term_planes[(2 * M + 2 * m + 1) * ld + i] == r[(i * M + m) * 2 + 1];
            ok = ok && h.obs[(i * 2 * M + m) * 2] == 1.0 && h.obs[(i * 2 * M + m) * 2 + 1] == 0.0; /* next episode: u = 1 */
        }
        bad += !ok;
        nconv += conv[i] != 0;
    }
    printf("envs %lld  converged %lld  mismatches %lld\n", (long long)N, (long long)nconv, (long long)bad);
    CK(sdcgym_pipe_destroy(pipe));
    return bad == 0 && nconv > 0 ? 0 : 1;
}
