#!/bin/bash
# round 2, call Y: one full ncu capture of the step's kernels on the final build (feeds profiles/traffic.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0 --cpu-enc-frames 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group|k_me_frac|k_rdoq|k_mc_batch|k_fwd_tq|k_inv_tq" --launch-skip 19 --launch-count 19 -f -o $O/r02y_prof $CMD > $O/r02y_ncu_f.log 2>&1; echo "ncu full rc=$?"
ls -la $O/r02y_prof.ncu-rep
