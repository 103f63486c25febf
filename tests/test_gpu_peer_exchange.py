"""GPU, two processes (ranks) sharing ONE device: the multi-rank VecNormalize whose moment all-reduce runs inside the
statistics kernel over CUDA IPC peer memory (csrc/vecnorm.cu update_kernel<.., DIST>, dist.PeerExchange), against the
collective path (accumulate -> all-reduce -> merge) and against a single process that normalises the whole batch.
The process group (gloo here: two ranks cannot share a GPU under NCCL) only carries the IPC handles at set-up."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_GLOBAL, M, STEPS = 6000, 5, 4
KW = dict(M=M, dt=1.0, restol=1e-10, seed=11, lambda_real_interval=[-100, 0], lambda_imag_interval=[-10, 0],
          reward_iteration_only=False)


def _actions(step):
    return np.random.default_rng(500 + step).uniform(-1, 1, (N_GLOBAL, M))


def _run(env, lo, hi):
    import sdc_gym_b200  # noqa: F401

    env.reset()
    outs = []
    for s in range(STEPS):
        a = torch.as_tensor(_actions(s)[lo:hi], device=env.venv.device)
        out = env.step_tensor(a)
        outs.append((out["reward"].clone(), out["obs_planes"].clone()))
    return outs


def _worker(rank, ws, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        import sdc_gym_b200
        from sdc_gym_b200 import dist as sdist

        lo, cnt = sdist.shard_range(N_GLOBAL, rank, ws)
        envs = {}
        for mode in ("peer", "nccl"):
            v = sdist.make_sharded("sdc-v1", N_GLOBAL, **KW)
            envs[mode] = sdc_gym_b200.VecNormalize(v, norm_obs=True, norm_reward=True, sync=True if mode == "peer" else "nccl")
        res = {m: _run(e, lo, lo + cnt) for m, e in envs.items()}
        # reset: observation statistics alone; every step: ONE exchange for the observation and the return statistics
        assert envs["peer"]._xchg_obs is not None and envs["peer"]._xchg_obs.seq == 1
        assert envs["peer"]._xchg_both.seq == STEPS and envs["peer"]._xchg_ret.seq == 0
        assert envs["nccl"]._xchg_obs is None
        for (r1, o1), (r2, o2) in zip(res["peer"], res["nccl"]):
            assert torch.equal(r1, r2) and torch.equal(o1, o2)
        stats = {}
        for m, e in envs.items():
            stats[m] = [t.cpu().numpy().copy() for t in (e.obs_rms.mean, e.obs_rms.var, e.obs_rms.count2,
                                                         e.ret_rms.mean, e.ret_rms.var, e.ret_rms.count2)]
        for a, b in zip(stats["peer"], stats["nccl"]):
            assert np.array_equal(a, b), "in-kernel exchange and collective path must give the same normaliser bits"
        gathered = [None] * ws
        dist.all_gather_object(gathered, [x.tolist() for x in stats["peer"]])
        assert gathered[0] == gathered[1], "normalisers must be identical on all ranks"
        np.save(os.path.join(tmp, f"stats{rank}.npy"), np.concatenate([x.ravel() for x in stats["peer"]]))
        for e in envs.values():
            if e._xchg_obs is not None:
                e._xchg_obs.close()
                e._xchg_ret.close()
                e._xchg_both.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_in_kernel_peer_exchange_matches_collective_and_single_process(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "stats0.npy")
    import sdc_gym_b200

    env = sdc_gym_b200.VecNormalize(sdc_gym_b200.make("sdc-v1", num_envs=N_GLOBAL, **KW), norm_obs=True, norm_reward=True)
    _run(env, 0, N_GLOBAL)
    ref = np.concatenate([t.cpu().numpy().ravel() for t in (env.obs_rms.mean, env.obs_rms.var, env.obs_rms.count2,
                                                            env.ret_rms.mean, env.ret_rms.var, env.ret_rms.count2)])
    # same statistics up to the summation order of the partial sums (two shards vs one)
    assert np.allclose(got, ref, rtol=1e-11, atol=1e-13)
