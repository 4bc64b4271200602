"""CPU: the reference arm of bench.py (`--impl reference`) honours the driver's contract -- one JSON line from rank 0 with the
arm's metric / unit / config, `impl: "reference"`, a `cpu_baseline` describing the run and a zero-copy `e2e`; other ranks
print nothing and exit 0.  (The CUDA arm needs a GPU; its line is checked by the GPU runs recorded under profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["metric"].split(" (")[0] in base["metric"]          # "pairs/s @540x960 D=192"
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 1.0) < 1e-6
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "576x960" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout[-200:], out.stderr[-300:])
