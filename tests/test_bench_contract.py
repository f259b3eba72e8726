"""bench.py's reference arm runs on the host cores only, so its JSON contract can be checked here (no GPU): one line, the keys the
driver reads, the same metric / unit / config as the GPU arm, zero copy bytes, a cpu_baseline that describes the run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhmref.so")) and not os.path.isdir("/root/reference/source"):
        import pytest
        pytest.skip("neither oracle/_ref/libhmref.so nor /root/reference: the reference arm has nothing to time")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-ctus-per-core", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"] == bench.workload_config() and "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "CTUs per step" in cb["sample"]


def test_workload_config_says_how_l2_is_handled():
    sys.path.insert(0, ROOT)
    import bench
    c = bench.workload_config()
    assert "no explicit flush" in c["l2"] and "126 MB L2" in c["l2"]
    assert "SAD tables" not in c["workload"]          # the round-2 form has none
