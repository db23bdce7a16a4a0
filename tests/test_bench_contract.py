"""bench.py's reference arm runs without a GPU: its JSON line must carry the keys the driver reads (the GPU arm's line is produced
on the B200 at round end; its extra keys are checked here only by name in the source)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "structured-layer fwd+bwd samples/sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("C5: SSS 4096->1000")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "batch 256" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_gpu_arm_emits_the_contract_keys():
    src = open(os.path.join(REPO, "bench.py")).read()
    for key in ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "h2d_bytes_per_step", "d2h_bytes_per_step", "traffic",
                "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "sm_max_mhz", "reasons"):
        assert key in src, key
