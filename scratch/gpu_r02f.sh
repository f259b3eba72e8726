#!/bin/bash
# round 2, call F: staged first sweep in the group kernel: parity + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02f_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02f_parity.log
tail -8 $O/r02f_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02f_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02f_props.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 5 --cpu-enc-frames 0 > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench rc=$?"; tail -3 $O/r02f_bench.err
