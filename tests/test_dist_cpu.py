"""CPU, world_size 2 over gloo: the N > 1 host logic - env sharding, statistics reduction and the cross-rank
merge of normaliser moments (the same arithmetic the device kernels run, restated in dist.chan_merge)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdc_gym_b200 import dist as sdist
from sdc_gym_b200 import rng as host_rng


def test_shard_range_partitions_the_index_space():
    for n in (0, 1, 7, 8, 1000, 2**20 + 3):
        for ws in (1, 2, 3, 8):
            parts = [sdist.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (o1, c1), (o2, _) in zip(parts, parts[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sdist.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n_global = 1001
        off, cnt = sdist.shard_range(n_global, *sdist.world())
        # lambda streams: a shard sees exactly its slice of the global stream
        lam_g = host_rng.lambda_stream(5, np.arange(n_global), 0, (-100, 0), (-10, 0))
        lam_l = host_rng.lambda_stream(5, np.arange(off, off + cnt), 0, (-100, 0), (-10, 0))
        assert np.array_equal(lam_l, lam_g[off:off + cnt])
        # rollout statistics
        rng = np.random.default_rng(0)
        reward = torch.as_tensor(rng.normal(size=n_global))
        niter = torch.as_tensor(rng.integers(1, 51, n_global).astype(np.int32))
        flags = torch.as_tensor(rng.integers(0, 8, n_global).astype(np.uint8))
        stats = sdist.RolloutStats("cpu")
        sl = slice(off, off + cnt)
        stats.update(dict(reward=reward[sl], niter=niter[sl], flags=flags[sl]))
        red = stats.reduce()
        done = (flags & 1) != 0
        assert red["env_steps"] == n_global and red["episodes"] == float(done.sum())
        assert abs(red["sum_reward"] - float(reward.sum())) < 1e-9
        assert red["sum_niter"] == float((niter.double() * done.double()).sum())
        assert red["converged"] == float((((flags & 2) != 0) & done).sum())
        # normaliser moments: all-reduced shifted sums + Chan merge == single-process merge of the whole batch
        P = 6
        mean, var, count = np.zeros(P), np.ones(P), 1e-4
        mean_s, var_s, count_s = mean.copy(), var.copy(), count
        for step in range(5):
            x = np.random.default_rng(100 + step).normal(loc=3.0 + step, scale=2.0, size=(P, n_global))
            xl = x[:, off:off + cnt]
            sums = torch.as_tensor(np.concatenate([(xl - mean[:, None]).sum(1), ((xl - mean[:, None]) ** 2).sum(1), [cnt]]))
            sdist.all_reduce_sum(sums)
            sums = sums.numpy()
            mean, var, count = sdist.chan_merge(mean, var, count, sums[:P], sums[P:2 * P], sums[2 * P])
            mean_s, var_s, count_s = sdist.chan_merge(mean_s, var_s, count_s, (x - mean_s[:, None]).sum(1),
                                                      ((x - mean_s[:, None]) ** 2).sum(1), n_global)
            # reference: SB3's update_from_moments with batch mean / population variance
        assert np.allclose(mean, mean_s, rtol=1e-12, atol=1e-12) and np.allclose(var, var_s, rtol=1e-12)
        assert count == count_s
        allx = np.concatenate([np.random.default_rng(100 + s).normal(loc=3.0 + s, scale=2.0, size=(P, n_global))
                               for s in range(5)], axis=1)
        assert np.allclose(mean, allx.mean(1), rtol=1e-6) and np.allclose(var, allx.var(1), rtol=1e-5)
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_chan_merge_matches_sb3_formula():
    rng = np.random.default_rng(1)
    mean, var, count = np.zeros(3), np.ones(3), 1e-4
    m2, v2, c2 = mean.copy(), var.copy(), count
    for _ in range(4):
        x = rng.normal(2.0, 3.0, size=(50, 3))
        mean, var, count = sdist.chan_merge(mean, var, count, (x - mean).sum(0), ((x - mean) ** 2).sum(0), 50)
        bm, bv, bc = x.mean(0), x.var(0), 50  # stable_baselines3.common.running_mean_std.update_from_moments
        delta = bm - m2
        tot = c2 + bc
        new_mean = m2 + delta * bc / tot
        M2 = v2 * c2 + bv * bc + np.square(delta) * c2 * bc / tot
        m2, v2, c2 = new_mean, M2 / tot, tot
    assert np.allclose(mean, m2, rtol=1e-13) and np.allclose(var, v2, rtol=1e-12) and count == c2
