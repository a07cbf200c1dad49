"""bench.py's driver contract that can be checked without a GPU: the reference arm prints
exactly one JSON line with the agreed keys, and the sbx arm fails loudly without CUDA."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                        capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
  res = _run("--impl", "reference", "--steps", "2", "--warmup", "3", "--cpu-sample-envs", "4")
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [l for l in res.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, res.stdout
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "building_env_steps_per_sec"
  assert d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
  assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
  assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
  assert "workload" in d["config"]


def test_reference_arm_non_zero_ranks_exit_silently():
  env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
  res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert res.returncode == 0 and res.stdout.strip() == ""


def test_sbx_arm_fails_loudly_without_cuda():
  import torch
  if torch.cuda.is_available():
    return
  res = _run("--steps", "2", "--warmup", "3")
  assert res.returncode != 0
  assert "no CPU fallback" in (res.stderr + res.stdout)
