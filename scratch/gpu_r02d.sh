#!/bin/bash
# round 2, call D: group search v3 (pair-based consume): parity, bi-pred refinement, bench line, ncu of the group kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02d_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02d_parity.log
tail -12 $O/r02d_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02d_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02d_props.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 5 --cpu-enc-frames 0 > $O/r02d_bench.json 2> $O/r02d_bench.err; echo "bench rc=$?"; tail -3 $O/r02d_bench.err
( timeout 900 python -m pytest tests/test_hm_md5.py -x -q -m gpu -k "test_bitstream_md5_identical_to_reference and (randomaccess or lowdelay_main)" -s > $O/r02d_md5.log 2>&1; echo "md5 rc=$?" >> $O/r02d_md5.log ) &
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group" --launch-skip 1 --launch-count 1 -f -o $O/r02d_prof $CMD > $O/r02d_ncu_f.log 2>&1; echo "ncu full rc=$?"
wait
tail -6 $O/r02d_md5.log
