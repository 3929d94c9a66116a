"""bench.py's reference arm runs on the CPU: check the JSON line it prints (the contract the driver parses)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "spgemm_gflops" and line["unit"] == "GFLOP/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # both arms run the same headline workload at every N (configs[3]); the reference arm on a stated sample of it
    assert "configs[3]" in line["config"]["workload"] and line["config"]["name"] == "er8m"
    assert set(line["config"]) == {"workload", "name", "scale_down"} and line["scaling"] == "strong"
    assert "1/64" in line["cpu_baseline"]["sample"]
