"""Full-size parity report (GPU + host cores): the CUDA path against the rounding-exact CPU oracle on EVERY env of
the benchmark workload (not a subsample).  The oracle runs on all host cores (it is test infrastructure; see
oracle/sdc_exact.c).  Prints one JSON line per configuration with mismatch counts; `> profiles/parity_report_*.jsonl`.

    python tools/parity_report.py [--envs 1048576]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np


def _oracle_chunk(args):
    kind, M, lam, act, prec, prec_type, steps = args
    from oracle import exact
    from sdc_gym_b200.collocation import collocation_matrix
    from sdc_gym_b200.precond import fixed_preconditioner
    Q = collocation_matrix(M)
    n = lam.shape[0]
    u, r = exact.reset(Q, 1.0, lam)
    rinit, niter = r.copy(), np.zeros(n, np.int32)
    Qd = fixed_preconditioner(prec, M, Q) if prec else None
    out = None
    for s in range(steps):
        out = exact.step(kind, Q, 1.0, lam, u, r, niter, rinit, None if prec else act[s], prec_type="fixed" if prec else prec_type,
                         Qd_fixed=Qd, do_scale=(prec_type == "diag"))
    return u, r, niter.copy(), out["resnorm"], out["reward"], out["done"], out["err"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--only-round2", action="store_true", help="only the cases added in round 2")
    args = ap.parse_args()
    import torch
    import sdc_gym_b200
    from sdc_gym_b200 import _lib
    from sdc_gym_b200.precond import fixed_preconditioner, num_actions
    cores = os.cpu_count() or 1
    pool = mp.get_context("spawn").Pool(cores)
    rng = np.random.default_rng(2026)
    cases = [("sdc-v0", 5, "diag", None, "uniform", args.envs, 1), ("sdc-v0", 5, "diag", None, "good", args.envs, 1),
             ("sdc-v0", 5, "diag", "LU", "-", args.envs // 4, 1), ("sdc-v0", 5, "lower_tri", None, "small", args.envs // 4, 1),
             ("sdc-v0", 7, "strictly_lower_tri", None, "small", args.envs // 8, 1),
             ("sdc-v0", 9, "lower_tri", None, "small", args.envs // 16, 1),
             ("sdc-v1", 5, "diag", None, "good", args.envs // 4, 8),
             # round 2: the phased dense solve (forced: phased=True) on the config-3 workload, and the diagonal full
             # solve with C in tensor memory (the default at M = 3, 4)
             ("sdc-v0", 5, "strictly_lower_tri", None, "cfg3", args.envs, 1, True),
             ("sdc-v0", 5, "lower_tri", None, "cfg3", args.envs // 2, 1, True),
             ("sdc-v0", 7, "strictly_lower_tri", None, "cfg3", args.envs // 4, 1, True),
             ("sdc-v0", 3, "lower_tri", None, "cfg3", args.envs // 2, 1, True),
             ("sdc-v0", 4, "diag", None, "uniform", args.envs, 1), ("sdc-v0", 3, "diag", None, "uniform", args.envs, 1)]
    if args.only_round2:
        cases = cases[7:]
    for case in cases:
        kind, M, pt, prec, mode, n, steps = case[:7]
        phased = case[7] if len(case) > 7 else None
        lam = rng.uniform(-100, 0, n) + 1j * rng.uniform(-10, 0, n)
        A = num_actions(M, pt)
        if prec:
            act = None
        elif mode == "uniform":
            act = rng.uniform(-1, 1, (steps, n, A))
        elif mode == "good":
            x = np.diag(fixed_preconditioner("min", M))
            act = 2 * (x[None, None] + rng.uniform(-0.03, 0.03, (steps, n, M))) - 1
        elif mode == "cfg3":
            act = rng.uniform(0, 0.3, (steps, n, A))
        else:
            act = rng.uniform(0, 0.12, (steps, n, A))
        env = sdc_gym_b200.make(kind, num_envs=n, M=M, dt=1.0, restol=1e-10, prec=prec, prec_type=pt,
                                do_scale=(pt == "diag"), blas_variant=_lib.BLAS_SKYLAKEX, autoreset=False, phased=phased,
                                lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0])
        env.reset(lam=lam)
        t0 = time.perf_counter()
        for s in range(steps):
            out = env.step_tensor(None if prec else torch.as_tensor(act[s], device=env.device))
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t0
        snap = env._snapshot()
        g = dict(u=snap["obs"][:, 0], r=snap["obs"][:, 1], niter=out["niter"].cpu().numpy(),
                 res=out["residual"].cpu().numpy(), rew=out["reward"].cpu().numpy(), flags=out["flags"].cpu().numpy())
        t0 = time.perf_counter()
        bounds = np.linspace(0, n, cores * 4 + 1).astype(int)
        jobs = [(kind, M, lam[a:b], None if prec else act[:, a:b], prec, pt, steps) for a, b in zip(bounds, bounds[1:]) if b > a]
        parts = pool.map(_oracle_chunk, jobs)
        t_cpu = time.perf_counter() - t0
        u, r, nit, res, rew, done, err = (np.concatenate([p[k] for p in parts]) for k in range(7))
        eq = lambda a, b: int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))
        conv_or_done = ((g["flags"] & 2) != 0) if kind == "sdc-v0" else ((g["flags"] & 1) != 0)
        print(json.dumps({
            "kind": kind, "M": M, "prec_type": pt if not prec else prec, "actions": mode, "envs": n, "steps": steps,
            "launch": ("phased" if getattr(env, "phased", False) else "single"),
            "handed_over": (int(env.phase_count[0]) if getattr(env, "phased", False) else None),
            "mismatch_niter": int(np.sum(g["niter"] != nit)), "mismatch_done_or_converged": int(np.sum(conv_or_done != done)),
            "mismatch_err": int(np.sum(((g["flags"] & 4) != 0) != err)), "mismatch_residual_norm": eq(g["res"], res),
            "mismatch_u_components": eq(g["u"].real, u.real) + eq(g["u"].imag, u.imag),
            "mismatch_r_components": eq(g["r"].real, r.real) + eq(g["r"].imag, r.imag),
            "max_rel_reward_diff": float(np.max(np.abs(g["rew"] - rew) / np.maximum(np.abs(rew), 1e-300))),
            "mean_niter": float(nit.mean()), "converged_frac": float(done.mean()), "err_frac": float(err.mean()),
            "gpu_seconds": round(t_gpu, 4), "oracle_seconds_all_cores": round(t_cpu, 2), "host_cores": cores}), flush=True)
        del env
    pool.close()


if __name__ == "__main__":
    main()
