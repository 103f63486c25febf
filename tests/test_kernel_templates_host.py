"""CPU: the kernel bodies themselves (sdc_gym_b200/csrc/*.cuh), compiled for the host by tests/host_shim,
replay every golden case bit-for-bit.  This is a pre-flight check of the CUDA code without a GPU; the same
replay runs against libsdcgym.so on the GPU in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from sdc_gym_b200.precond import num_actions
from tests import host_shim
from tests.helpers import case_ids
from tests.replay import replay_case


class ShimBackend:
    def __init__(self, meta, g):
        d = host_shim.make_desc(meta["kind"], meta["M"], prec=meta["prec"], prec_type=meta["prec_type"] if meta["prec"] is None else "diag",
                                dt=meta["dt"], restol=meta["restol"], cplx=meta["cplx"], do_scale=meta["do_scale"],
                                use_doubles=meta.get("use_doubles", True),
                                strategy=meta["strategy"], step_penalty=meta["step_penalty"],
                                residual_weight=meta["residual_weight"], norm_factor=meta["norm_factor"], Q=g["Q"])
        self.b = host_shim.ShimBatch(d, meta["n"], collect=meta["collect"])
        self.n_act = 0 if meta["prec"] is not None else num_actions(meta["M"], meta["prec_type"])

    def reset(self, lam):
        return self.b.reset(lam)

    def step(self, actions):
        return self.b.step(actions)

    def old_states_host(self):
        return self.b.old_states


@pytest.mark.parametrize("name", case_ids())
def test_kernel_templates_replay_golden(name):
    replay_case(name, ShimBackend)
