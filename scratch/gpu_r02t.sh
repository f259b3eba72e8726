#!/bin/bash
# round 2, call M: what the driver runs at round end -- the whole -m gpu suite, smoke(), the reference arm and the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > $O/r02t_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -6 $O/r02t_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02t_smoke.log
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 ) > $O/r02t_bench_reference.json 2> $O/r02t_bench_reference.err; echo "reference arm rc=$?"
( time timeout 1500 python bench.py ) > $O/r02t_bench.json 2> $O/r02t_bench.err; echo "bench rc=$?"; tail -4 $O/r02t_bench.err
