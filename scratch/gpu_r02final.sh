#!/bin/bash
# round 2, final call: short bench line + launch list of the final build, then the whole -m gpu suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --hm-frames 0 --cpu-enc-frames 0 > $O/r02f_bench_short.json 2> $O/r02f_bench_short.err; echo "bench rc=$?"; tail -2 $O/r02f_bench_short.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0 --cpu-enc-frames 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches.csv $CMD > $O/r02f_ncu_l.log 2>&1; echo "ncu list rc=$?"
( time timeout 1150 python -m pytest tests/ -x -q -m gpu ) > $O/r02f_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -6 $O/r02f_gpu_suite.log
