"""N > 1 host logic on CPU (gloo, world size 2): independent sequences per rank, max-over-ranks timing,
rank 0 alone reports; and the --impl reference arm under torchrun (rank 0 works, other ranks exit 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


def test_rank_sharding_and_timing_gloo():
    r = _torchrun([os.path.join(ROOT, "tests", "mp_worker.py")], 29611)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["world"] == 2 and len(set(d["seeds"])) == 2          # independent sequences
    assert d["max_ms"] == 11.0                                     # max over ranks, not rank 0's own time
    assert abs(d["fps"] - 2 * 1e3 / 11.0) < 1e-9                   # whole-job aggregate: N pictures per max step time


def test_reference_arm_under_torchrun_prints_one_line():
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--ref-ctus-per-core", "1"], 29612)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] in ("reference", "port") and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0
