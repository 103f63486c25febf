"""CPU: host-side logic that needs no GPU - collocation identities, library symbols, RNG restatement,
info containers, argument validation through the C ABI."""
import ctypes

import numpy as np
import pytest

from sdc_gym_b200 import _lib, collocation_matrix, precond, rng
from sdc_gym_b200.collocation import CollGaussRadauRight


@pytest.mark.parametrize("M", range(2, 10))
def test_collocation_identities(M):
    c = CollGaussRadauRight(M, 0, 1)
    Q, t = c.Qmat[1:, 1:], c.nodes
    ulp = np.finfo(float).eps
    assert np.all(np.abs(Q @ np.ones(M) - t) <= 2 * ulp)
    assert np.all(np.abs(Q @ t - t * t / 2) <= 2 * ulp)
    assert abs(Q[-1, -1] - 1 / M**2) <= ulp
    assert abs(Q[-1].sum() - 1) <= 2 * ulp and t[-1] == 1.0
    assert np.all(np.diff(t) > 0) and c.Qmat[0].sum() == 0 and c.Qmat[:, 0].sum() == 0
    assert np.allclose(np.cumsum(c.delta_m), t)


def test_collocation_known_values_radau_iia_3():
    Q = collocation_matrix(3)
    s6 = np.sqrt(6.0)
    ref = np.array([[(88 - 7 * s6) / 360, (296 - 169 * s6) / 1800, (-2 + 3 * s6) / 225],
                    [(296 + 169 * s6) / 1800, (88 + 7 * s6) / 360, (-2 - 3 * s6) / 225],
                    [(16 - s6) / 36, (16 + s6) / 36, 1 / 9]])
    assert np.allclose(Q, ref, rtol=0, atol=2e-16)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _lib.exported_symbols()
    assert {"sdcgym_reset", "sdcgym_step", "sdcgym_export_obs", "sdcgym_spectral_radius"} <= set(names)
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/sdcgym.h but not exported"
    assert L.sdcgym_abi_version() == _lib.ABI_VERSION


def test_num_actions_and_support_matrix():
    L = _lib.load()
    for M in range(2, 10):
        for pt, code in _lib.PREC_TYPES.items():
            assert L.sdcgym_supported(M, code) == 1
            if pt != "fixed":
                assert L.sdcgym_num_actions(M, code) == precond.num_actions(M, pt)
    assert L.sdcgym_supported(1, 0) == 0 and L.sdcgym_supported(10, 0) == 0 and L.sdcgym_supported(5, 9) == 0
    assert L.sdcgym_num_actions(5, 17) == -1


def test_argument_errors_do_not_touch_the_gpu():
    L = _lib.load()
    d, st, io = _lib.EnvDesc(), _lib.State(), _lib.StepIO()
    d.M, d.prec_type, d.max_iters = 5, 0, 50
    assert L.sdcgym_step(None, ctypes.byref(st), ctypes.byref(io), None) == -3
    d.M = 12
    assert L.sdcgym_step(ctypes.byref(d), ctypes.byref(st), ctypes.byref(io), None) == -2
    d.M, d.env_kind = 5, 7
    assert L.sdcgym_step(ctypes.byref(d), ctypes.byref(st), ctypes.byref(io), None) == -1
    d.env_kind, st.N, st.ld = 0, 4, 2
    assert L.sdcgym_reset(ctypes.byref(d), ctypes.byref(st), None, None, None, None) == -1
    st.N, st.ld = 4, 4
    assert L.sdcgym_reset(ctypes.byref(d), ctypes.byref(st), None, None, None, None) == -3  # null planes
    st.N = 0
    assert L.sdcgym_reset(ctypes.byref(d), ctypes.byref(st), None, None, None, None) == 0  # empty batch: no-op
    assert L.sdcgym_step(ctypes.byref(d), ctypes.byref(st), ctypes.byref(io), None) == 0
    assert L.sdcgym_export_obs(5, 0, 0, None, None, None) == 0


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    z = rng.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    z = rng.philox4x32_10(f, f, f, f, f, f)
    assert [int(v) for v in z] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    z = rng.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(v) for v in z] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_lambda_stream_in_box_and_shard_invariant():
    lam = rng.lambda_stream(3, np.arange(1000), 0, (-100, 0), (-10, 0))
    assert np.all((lam.real >= -100) & (lam.real < 0) & (lam.imag >= -10) & (lam.imag < 0))
    lam2 = rng.lambda_stream(3, np.arange(500, 1000), 0, (-100, 0), (-10, 0))
    assert np.array_equal(lam[500:], lam2)
    assert not np.array_equal(lam, rng.lambda_stream(4, np.arange(1000), 0, (-100, 0), (-10, 0)))
    assert not np.array_equal(lam, rng.lambda_stream(3, np.arange(1000), 1, (-100, 0), (-10, 0)))


def test_lazy_infos_protocol():
    from sdc_gym_b200.vec_env import LazyInfos
    done = np.array([True, False, True])
    term = np.arange(3 * 2 * 2).reshape(3, 2, 2).astype(np.complex128)
    calls = []
    infos = LazyInfos(np.array([50, 3, 7]), np.array([1.0, 2.0, 3e-11]), np.array([1j, 2j, 3j]), done,
                      np.array([True, False, False]), lambda: calls.append(1) or term)
    assert len(infos) == 3 and not calls
    assert set(infos[1]) == {"residual", "niter", "lam"} and not calls
    assert infos[0]["TimeLimit.truncated"] is False and np.array_equal(infos[0]["terminal_observation"], term[0])
    assert "TimeLimit.truncated" not in infos[2] and infos[2]["niter"] == 7 and infos[-1]["lam"] == 3j
    assert len(calls) == 1
    assert [i["niter"] for i in infos] == [50, 3, 7]
    with pytest.raises(IndexError):
        infos[3]


def test_env_construction_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import sdc_gym_b200
    with pytest.raises(_lib.SdcGymError):
        sdc_gym_b200.make("sdc-v0", num_envs=4, M=3, dt=1.0, restol=1e-10)
    with pytest.raises(KeyError):
        sdc_gym_b200.make("sdc-v9", num_envs=4, M=3, dt=1.0, restol=1e-10)


def test_block_layout_offsets_are_aligned_and_disjoint():
    L = _lib.load()
    lay = _lib.BlockLayout()
    assert L.sdcgym_block_layout_init(5, 1000, ctypes.byref(lay)) == 0
    sizes = dict(obs_u=1000 * 80, reward=8000, residual=8000, lam=16000, niter=4000, flags=1000, obs_r=1000 * 80)
    order = ["obs_u", "reward", "residual", "lam", "niter", "flags", "obs_r"]
    end = 0
    for k in order:
        off = getattr(lay, k)
        assert off % 256 == 0 and off >= end, k
        end = off + sizes[k]
    assert lay.total >= end and lay.total % 256 == 0 and lay.N == 1000 and lay.M == 5
    # skip_u transfers the contiguous tail [reward, total): everything but the u rows
    assert lay.obs_u == 0 and lay.obs_r > lay.flags
    assert L.sdcgym_block_layout_init(5, 0, ctypes.byref(lay)) == 0 and lay.total == 0
    assert L.sdcgym_block_layout_init(12, 4, ctypes.byref(lay)) == -1
    assert L.sdcgym_block_layout_init(5, 4, None) == -3
    assert L.sdcgym_pipe_step_block(None, None, None, None, None, None, None) == -3


def test_host_result_set_is_reused_only_when_the_caller_let_go():
    """ownership without copies: a result block is written again only when no array handed out from it (nor any view
    derived from one) is referenced outside the env (vec_env._HostSet)"""
    import torch
    from sdc_gym_b200.vec_env import LazyInfos, _HostSet

    class FakeTorch:  # page-locked allocation needs CUDA; the logic under test does not
        uint8 = torch.uint8

        @staticmethod
        def zeros(n, dtype, pin_memory):
            return torch.zeros(n, dtype=dtype)

    L = _lib.load()
    lay = _lib.BlockLayout()
    N, M = 37, 5
    assert L.sdcgym_block_layout_init(M, N, ctypes.byref(lay)) == 0
    hs = _HostSet(FakeTorch, lay, N, M, True)
    assert hs.free() and hs.obs.shape == (N, 2, M) and hs.obs.dtype == np.complex128
    assert np.all(hs.obs[:, 0] == 1.0) and np.all(hs.obs[:, 1] == 0.0)
    hs.root[lay.obs_r: lay.obs_r + 16].view(np.complex128)[0] = 2 + 3j  # what the DMA writes
    assert hs.obs[0, 1, 0] == 2 + 3j and hs.free()
    for make_ref in (lambda: hs.obs, lambda: hs.obs[3], lambda: hs.reward[2:4], lambda: hs.lam.real,
                     lambda: LazyInfos(hs.niter, hs.residual, hs.lam, np.ones(N, bool), None, None),
                     lambda: torch.from_numpy(hs.obs), lambda: hs.flags.view(np.bool_), lambda: hs.dones,
                     lambda: hs.dones[1:3]):
        ref = make_ref()
        assert not hs.free()
        del ref
        assert hs.free()
    copy = np.ascontiguousarray(hs.obs)  # a real copy does not pin the block
    assert hs.free() and copy.flags.c_contiguous
    hs0 = _HostSet(FakeTorch, lay, N, M, False)
    assert np.all(hs0.obs == 0)


def test_vecnormalize_load_rejects_foreign_files(tmp_path):
    import pickle

    from sdc_gym_b200.vec_normalize import VecNormalize

    p = tmp_path / "vecnormalize.pkl"
    with open(p, "wb") as f:
        pickle.dump({"obs_rms": "whatever an SB3 pickle holds"}, f)
    with pytest.raises(ValueError, match="stable-baselines"):
        VecNormalize.load(str(p), None)  # rejected before anything is unpickled or any device is touched
    q = tmp_path / "other.npz"
    np.savez(q, a=np.zeros(3))
    with pytest.raises(ValueError, match="not a sdc_gym_b200"):
        VecNormalize.load(str(q), None)


def test_register_gym_is_a_noop_without_gym():
    import importlib.util

    import sdc_gym_b200

    ids = sdc_gym_b200.register_gym()
    if importlib.util.find_spec("gym") is None and importlib.util.find_spec("gymnasium") is None:
        assert ids == []
    else:
        assert all(i.split(":")[1] in sdc_gym_b200.REGISTRY for i in ids)
    from sdc_gym_b200 import gym_adapter

    assert callable(gym_adapter.make_single)
