"""bench.py's driver contract, checked on the arm that runs without a GPU: `--impl reference` prints ONE JSON line with
the keys the driver parses (metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling /
vs_baseline / dtype / data / config.workload, plus impl, cpu_baseline and an e2e object without copies); the native arm
must fail loudly without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
  e = dict(os.environ)
  e.update(env or {})
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                        cwd=ROOT, env=e, timeout=600)


def test_reference_arm_line_follows_the_contract():
  r = _run("--impl", "reference", "--workload", "fill", "--gpus", "1", "--steps", "2", "--warmup", "1")
  assert r.returncode == 0, r.stderr[-800:]
  lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1, r.stdout[-800:]
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
  assert d["unit"] == "voxels/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
  assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
  assert d["value"] > 0 and abs(d["value"] - 2 * 128 ** 3 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
  cb = d["cpu_baseline"]
  assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
  assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
  r = _run("--impl", "reference", "--workload", "fill", "--gpus", "2", "--steps", "1", "--warmup", "0",
           env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
  assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_native_arm_needs_a_gpu():
  import torch
  if torch.cuda.is_available():
    import pytest
    pytest.skip("GPU present")
  r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
  assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
