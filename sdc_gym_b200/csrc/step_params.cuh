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

// Register-residency policy for C and occupancy target (measured on B200, profiles/tuning_r01.md):
// diag kernels keep Re(C) in registers and re-derive Im(C) = -zi*q from the constant bank (166 registers,
// 3 blocks of 128 threads per SM) up to M=5, Re only up to M=7, nothing beyond; dense kernels spend their
// registers on the M x M inverse instead.
template <int M>
struct HoldPolicy {
    static constexpr int diag = (M <= 7) ? 1 : 0;
    static constexpr int diag_minb = (M <= 5) ? 3 : 2;
    static constexpr int dense = (M <= 3) ? 2 : 0;
    static constexpr int dense_minb = 2;
    // sdc-v1 runs a single sweep per launch and is memory bound: nothing is worth holding, occupancy is
    static constexpr int step = 0;
    static constexpr int step_minb = (M <= 5) ? 4 : 2;
};

}  // namespace sdcgym
