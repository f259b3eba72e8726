#!/bin/bash
# round 2, call C: request-based group search: parity, bench line (short encoder leg), ncu of the group kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02c_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02c_parity.log
tail -12 $O/r02c_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02c_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02c_props.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02c_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 5 --cpu-enc-frames 0 > $O/r02c_bench.json 2> $O/r02c_bench.err; echo "bench rc=$?"; tail -3 $O/r02c_bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group" --launch-skip 1 --launch-count 1 -f -o $O/r02c_prof $CMD > $O/r02c_ncu_f.log 2>&1; echo "ncu full rc=$?"
