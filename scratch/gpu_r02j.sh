#!/bin/bash
# round 2, call J: CU-level fractional search with tile masks: parity, bench, launch list, full ncu captures of the step's kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rdoq.py -x -q -m gpu -k "me_ or rdoq_recon" > $O/r02j_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02j_parity.log
tail -8 $O/r02j_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02j_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02j_props.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu > $O/r02j_bench.json 2> $O/r02j_bench.err; echo "bench rc=$?"; tail -3 $O/r02j_bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02j_launches.csv $CMD > $O/r02j_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group|k_me_frac|k_rdoq|k_mc_batch|k_fwd_tq|k_inv_tq" --launch-skip 14 --launch-count 16 -f -o $O/r02j_prof $CMD > $O/r02j_ncu_f.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intra_rough" --launch-skip 2 --launch-count 2 -f -o $O/r02j_prof_intra $CMD > $O/r02j_ncu_i.log 2>&1; echo "ncu intra rc=$?"
