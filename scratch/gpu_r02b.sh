#!/bin/bash
# round 2, call B: parity of the group search (both forms), sanitizer pass, bench line, ncu launch list + full captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02b_parity.log 2>&1; echo "parity rc=$?" | tee -a $O/r02b_parity.log
tail -25 $O/r02b_parity.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_me_frame_prepass and group" > $O/r02b_sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a $O/r02b_sanitizer.log
tail -5 $O/r02b_sanitizer.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02b_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02b_props.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02b_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench rc=$?"; tail -3 $O/r02b_bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches.csv $CMD > $O/r02b_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_me_group|k_me_frac|k_rdoq" --launch-skip 5 --launch-count 6 -f -o $O/r02b_prof $CMD > $O/r02b_ncu_f.log 2>&1; echo "ncu full rc=$?"
