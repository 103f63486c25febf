"""GPU: the C ABI driven from a plain C program (no Python, no torch in the process): tests/c_abi/abi_roundtrip.c
links libsdcgym.so + libcudart and the CPU oracle, steps 20 000 sdc-v0 envs through sdcgym_pipe_step with host
arrays and checks bit-exact parity.  Also pins that the header compiles as C."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from sdc_gym_b200.collocation import collocation_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp):
    exe = os.path.join(tmp, "abi_roundtrip")
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(CUDA, "include"), os.path.join(ROOT, "tests", "c_abi", "abi_roundtrip.c"),
           os.path.join(ROOT, "oracle", "sdc_exact.c"), "-L", os.path.join(ROOT, "sdc_gym_b200"), "-lsdcgym",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "sdc_gym_b200"),
           "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "sdcgym.h"\nint main(void) { sdcgym_env_desc d; (void)d; return SDCGYM_ABI_VERSION - 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "hdr.o")])


@pytest.mark.gpu
def test_c_program_steps_envs_through_the_abi(tmp_path):
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA headers not available")
    exe = _build(str(tmp_path))
    qfile = tmp_path / "Q5.bin"
    collocation_matrix(5).astype(np.float64).tofile(qfile)
    out = subprocess.run([exe, "20000", str(qfile)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
