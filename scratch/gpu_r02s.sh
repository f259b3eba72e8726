#!/bin/bash
# round 2, call I: 16x16 CUs with their 8x8 children in the CU-level fractional search: parity + bench + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02s_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02s_parity.log
tail -8 $O/r02s_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02s_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02s_props.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu > $O/r02s_bench.json 2> $O/r02s_bench.err; echo "bench rc=$?"; tail -3 $O/r02s_bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02s_launches.csv $CMD > $O/r02s_ncu_l.log 2>&1; echo "ncu list rc=$?"
