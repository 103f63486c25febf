"""CPU: the `bench.py --impl reference` arm prints exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, SDCGYM_BENCH_REF_SECONDS="0.2", SDCGYM_BENCH_REF_CORES="2", OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "3"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 3 and d["vs_baseline"] is None
    from oracle import ref_loader
    want = "reference" if ref_loader.reference_path() is not None else "port"  # the unmodified env when it is reachable
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("SDC env-steps/sec")
    # ranks other than 0 do no work and print nothing
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True,
                         text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""
