#!/bin/bash
# round 2, call E: staged-window group search (warp per PU): parity, bi-pred refinement, packed results, bench, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rdoq.py -x -q -m gpu -k "me_ or rdoq_recon" > $O/r02e_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02e_parity.log
tail -12 $O/r02e_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02e_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02e_props.log
( timeout 900 python -m pytest tests/test_hm_md5.py -x -q -m gpu -k "test_bitstream_md5_identical_to_reference and (randomaccess or lowdelay_main)" -s > $O/r02e_md5.log 2>&1; echo "md5 rc=$?" >> $O/r02e_md5.log ) &
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 5 --cpu-enc-frames 0 > $O/r02e_bench.json 2> $O/r02e_bench.err; echo "bench rc=$?"; tail -3 $O/r02e_bench.err
TVC_GROUP_MINB=3 timeout 600 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu > $O/r02e_bench_minb3.json 2> $O/r02e_bench_minb3.err; echo "bench3 rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group" --launch-skip 1 --launch-count 1 -f -o $O/r02e_prof $CMD > $O/r02e_ncu_f.log 2>&1; echo "ncu full rc=$?"
wait
tail -6 $O/r02e_md5.log
