"""bench.py's reference arm runs on the CPU (the oracle on the host cores) and must keep the driver's JSON contract: one
line, the metric/unit of the GPU arm, `impl`, `cpu_baseline`, `e2e`; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    env["CUDA_VISIBLE_DEVICES"] = ""              # the arm must not need a GPU
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--small", "--steps", "1",
                           "--warmup", "0"], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    out = run()
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iters/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ANLS iters/sec")
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000.0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    out = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == ""
