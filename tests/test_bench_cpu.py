"""bench.py's reference arm (the reference algorithm on the host cores) runs without a GPU: check its JSON contract here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample", "4", "--workload", "vqa2_b64"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train questions/sec (fwd+bwd)" and d["unit"] == "questions/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and d["steps"] == 1
    # "reference": the unmodified reference from the git-ignored install baseline/_ref; "port": the oracle restatement when it is absent
    installed = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "sparse_graph_model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if installed else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_zero_only():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
