"""bench.py's reference arm runs on the CPU: check the JSON contract of the line it prints (the device arm prints the
same keys plus roofline / clocks and is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, PFG_BENCH_CPU_BUDGET_S="0.05")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["unit"] == "elements/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["vs_baseline"] is None and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the CPU arm runs a bounded sample of the named workload: its size is stated in the line
    assert line["config"]["cpu_sample_elements"] > 0 and line["scaling"] == "strong"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, PFG_BENCH_CPU_BUDGET_S="0.05", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
