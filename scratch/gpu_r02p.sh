#!/bin/bash
# round 2, call P: the default bench line of the final build, its launch list and one full ncu capture of the step's kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time timeout 1500 python bench.py ) > $O/r02p_bench.json 2> $O/r02p_bench.err; echo "bench rc=$?"; tail -4 $O/r02p_bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02p_launches.csv $CMD > $O/r02p_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group|k_me_frac|k_rdoq|k_mc_batch|k_fwd_tq|k_inv_tq" --launch-skip 14 --launch-count 14 -f -o $O/r02p_prof $CMD > $O/r02p_ncu_f.log 2>&1; echo "ncu full rc=$?"
ls -la $O/r02p_prof.ncu-rep
