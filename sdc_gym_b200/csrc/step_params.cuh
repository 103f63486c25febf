// step_params.cuh - translation of the C ABI structs (include/sdcgym.h) into the kernel parameter block.
#pragma once
#include "step_kernels.cuh"

namespace sdcgym {

template <int M>
inline void fill_params(StepParams<M>& p, const sdcgym_env_desc* d, const sdcgym_state* st) {
    for (int k = 0; k < M * M; k++) {
        p.Q[k] = d->Q[k];
        p.Qd[k] = d->Qd_fixed[k];
    }
    p.N = st->N;
    p.ld = st->ld;
    p.lam = st->lam;
    p.S = st->S;
    p.resnorm = st->resnorm;
    p.niter = st->niter;
    p.episodes = st->episodes;
    p.rng_ctr = st->rng_ctr;
    p.norm_init = st->norm_init;
    p.action = nullptr;
    p.a_es = p.a_cs = 0;
    p.reward = nullptr;
    p.flags = nullptr;
    p.info_res = nullptr;
    p.info_niter = nullptr;
    p.info_lam = nullptr;
    p.term = nullptr;
    p.old_states = nullptr;
    p.lam_in = nullptr;
    p.mask = nullptr;
    p.dt = d->dt;
    p.restol = d->restol;
    p.step_penalty = d->step_penalty;
    p.residual_weight = d->residual_weight;
    p.norm_factor = d->norm_factor;
    p.re_lo = d->lam_re_lo;
    p.re_hi = d->lam_re_hi;
    p.im_lo = d->lam_im_lo;
    p.im_hi = d->lam_im_hi;
    p.ix0 = d->interp_x0;
    p.ix1 = d->interp_x1;
    p.seed = d->seed;
    p.env_offset = d->env_offset;
    p.prec_type = d->prec_type;
    p.is_complex = d->action_is_complex;
    p.do_scale = d->do_scale;
    p.max_iters = d->max_iters;
    p.strategy = d->reward_strategy;
    p.autoreset = d->autoreset;
    p.curriculum = d->curriculum;
    p.log_restol_nf = log(d->restol * d->norm_factor);
    p.it_stop = 0x7fffffff;
    p.min_lanes = 0;
    // work buffers of the phased dense solve (launch_step decides whether to use them; single-launch kernels ignore them)
    p.cont_list = st->phase_list;
    p.cont_count = st->phase_count;
    p.pinv_scratch = st->phase_pinv;
}


template <int M>
inline void fill_step_io(StepParams<M>& p, const sdcgym_step_io* io) {
    p.action = io->action;
    p.a_es = io->action_env_stride;
    p.a_cs = io->action_comp_stride;
    p.reward = io->reward;
    p.flags = io->flags;
    p.info_res = io->info_residual;
    p.info_niter = io->info_niter;
    p.info_lam = io->info_lam;
    p.term = io->terminal_obs;
    p.old_states = io->old_states;
}

// Where the system matrix C lives and how many blocks share an SM (measured on B200, profiles/README.md):
//  * full-solve diag kernels, M >= 5: all of C in shared memory (HOLD 4; [2 M^2][block] doubles, conflict free).
//    The sweep is bound by dependent-issue latency, not FP64 throughput, so what pays is warps per SM: with C out
//    of the register file the M=5 kernel fits 16 warps/SM (8 % faster than Re(C) in registers at 12 warps/SM).
//    M = 7 / 9 do not fit C in registers at all (255 + spills before).
//  * M <= 4: C is small enough to stay in registers.
//  * dense kernels spend their registers on the M x M inverse; sdc-v1 runs one sweep per launch and is memory
//    bound: nothing is held, occupancy is maximised.
#ifndef SDCGYM_DIAG_MID
#define SDCGYM_DIAG_MID 7  // residency of the M = 5..7 diagonal kernels (4 or 7)
#endif
#ifndef SDCGYM_DENSE_SMALL
#define SDCGYM_DENSE_SMALL 7  // residency of the M = 4, 5 dense kernels (4 or 7; 7 measured 2-4 % faster)
#endif
#ifndef SDCGYM_DENSE_MID
#define SDCGYM_DENSE_MID 9  // residency of the M = 6, 7 dense kernels (0 or 9)
#endif
#ifndef SDCGYM_DENSE_BIG
#define SDCGYM_DENSE_BIG 9  // residency of the M = 8, 9 dense kernels (8 or 9, see step_kernels.cuh)
#endif
#ifndef SDCGYM_DENSE_MINB
#define SDCGYM_DENSE_MINB 3
#endif
template <int M>
struct HoldPolicy {
    // M = 5..7: C in shared memory as (re, im) pairs read with one LDS.128 (HOLD 7; planar HOLD 4 was 2.8 % slower);
    // M >= 8: 2 blocks of 64 threads lose to recomputing
    static constexpr int diag = (M <= 4) ? 1 : ((M <= 7) ? SDCGYM_DIAG_MID : 0);
    static constexpr int diag_minb = (M <= 5) ? 4 : 2;
    static constexpr int diag_block = 128;
    // dense kernels: M <= 5 keep the inverse in registers and C in shared memory (HOLD 4); M = 6, 7 run the inverse on
    // register-resident LU factors and re-derive C (HOLD 0); M = 8, 9 keep the LU work matrix, then Pinv, in shared
    // memory and re-derive C (HOLD 9; HOLD 8 = C instead of Pinv there measured 25 % slower) in 64-thread blocks.  (Keeping Pinv in shared memory as well - HOLD 5, parity-tested - leaves
    // only 2-4 warps per SM and measured slower than letting Pinv spill to local memory.)
    static constexpr int dense = (M <= 3) ? 2 : ((M <= 5) ? SDCGYM_DENSE_SMALL : ((M <= 7) ? SDCGYM_DENSE_MID : SDCGYM_DENSE_BIG));
    static constexpr int dense_block = (M <= 7) ? 128 : 64;
    static constexpr int dense_minb = (M <= 5) ? SDCGYM_DENSE_MINB : 2;  // 168 registers: 3 blocks/SM (170 would round up to 176 = 2 blocks)
    static constexpr int step = 0;
    static constexpr int step_minb = (M <= 5) ? 6 : ((M <= 7) ? 3 : 2);  // M=5: 80 regs, 24 warps/SM: +12 % (profiles/tune_r01_v1.log)
};

}  // namespace sdcgym
