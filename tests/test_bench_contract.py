"""bench.py contract on CPU: the reference arm prints exactly one JSON line with the keys the driver reads, and the b200 arm
refuses to run without a CUDA device (there is no CPU fallback of the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
  res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--inputs", "4")
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [l for l in res.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  line = json.loads(lines[0])
  for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
    assert key in line, key
  assert line["impl"] == "reference" and line["metric"] == "mm_gp_predict_states_per_s" and line["dtype"] == "f64"
  assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
  assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
  assert "workload" in line["config"] and line["vs_baseline"] is None


def test_b200_arm_fails_loudly_without_a_gpu():
  import torch
  if torch.cuda.is_available():
    return
  res = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-pathwise", "--no-policy-opt")
  assert res.returncode != 0
  assert "no CPU implementation" in res.stderr
