"""Times the VecNormalize kernels alone (CUDA events, 2^20 / 2^22 envs): the fused statistics launch (observation planes
+ return plane), the three-kernel sequence, and the apply pass.  One JSON object per line.
SDCGYM_LIB=<experiment build> selects another library."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdc_gym_b200

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="+", default=[1 << 20, 1 << 22])
ap.add_argument("--M", type=int, default=5)
ap.add_argument("--tag", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
KW = dict(dt=1.0, restol=1e-10, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0], seed=0)


def timed(fn, steps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3  # us


for N in args.envs:
    M = args.M
    env = sdc_gym_b200.make("sdc-v1", num_envs=N, M=M, reward_iteration_only=False, **KW)
    vn = sdc_gym_b200.VecNormalize(env, norm_obs=True, norm_reward=True)
    vn.reset()
    a = torch.rand((N, M), dtype=torch.float64, device=dev) * 2 - 1
    vn.step_tensor(a)
    P = 4 * M
    with torch.cuda.device(dev):
        both = timed(lambda: vn._update_both(N))
        obs = timed(lambda: vn._update(vn.obs_rms, env.S.data_ptr(), P, N, env.ld, vn._sums))
        r = vn.ret_rms
        import ctypes
        L = vn._L
        rets = timed(lambda: L.sdcgym_vecnorm_update_returns(
            N, env.reward.data_ptr(), vn.gamma, vn.returns.data_ptr(), r.mean.data_ptr(), r.var.data_ptr(),
            r.count2.data_ptr(), vn._rscratch.data_ptr(), vn._rsums.data_ptr(), vn._stream()))
        vn.fused_update = False
        three = timed(lambda: vn._update(vn.obs_rms, env.S.data_ptr(), P, N, env.ld, vn._sums))
        vn.fused_update = True
        app = timed(lambda: vn._normalize_planes(env.S, vn.norm_planes))
        step = timed(lambda: env.step_tensor(a))
        nstep = timed(lambda: vn.step_tensor(a))
    print(json.dumps({"tag": args.tag, "envs": N, "M": M,
                      "update_both_us": both, "update_both_GBps": (P + 3) * 8 * N / both / 1e3,
                      "update_obs_us": obs, "update_returns_us": rets, "three_kernel_obs_us": three,
                      "apply_us": app, "apply_GBps": 2 * P * 8 * N / app / 1e3,
                      "step_us": step, "normalised_step_us": nstep}), flush=True)
    del vn, env
