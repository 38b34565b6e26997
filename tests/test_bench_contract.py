"""bench.py's reference arm (`--impl reference`, the CPU leg the driver times next to the GPU arm) prints ONE JSON line
with the contract's keys, on this machine without a GPU.  The rotated-IoU workload is used because its reference arm
(oracle/_ref or the C port) finishes in seconds."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "riou",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "rotated-IoU Mpairs/sec" and d["unit"] == "Mpairs/s" and "workload" in d["config"]
    assert set(("kind", "cores", "sample", "value")) <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["steps"] >= 1          # the steps actually run
