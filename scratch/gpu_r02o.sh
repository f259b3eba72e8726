#!/bin/bash
# round 2, call O: clean timings (nothing else on the GPU): CU-level first search on / off, encoder leg with / without look-ahead
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
TVC_GROUP_CU=0 timeout 600 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu > $O/r02o_bench_cu0.json 2> $O/r02o_bench_cu0.err; echo "bench cu0 rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02o_bench.json 2> $O/r02o_bench.err; echo "bench rc=$?"; tail -3 $O/r02o_bench.err
TVC_BENCH_HM_HOOKS=me,frac,tables,frame,candgrid,nospec timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu > $O/r02o_bench_nospec.json 2> $O/r02o_bench_nospec.err; echo "bench nospec rc=$?"
python - <<'P'
import json
for f in ("gpurun_out/r02o_bench_cu0.json", "gpurun_out/r02o_bench.json", "gpurun_out/r02o_bench_nospec.json"):
    b = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, b["ms_per_step"], b["detail"]["phase_ms_per_step"], b["e2e"]["ms_per_step"])
    h = b["detail"].get("hm_encode")
    if h: print({k: h[k] for k in h if k != "hooks"}); print("\n".join(h["hooks"]))
P
