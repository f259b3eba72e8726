#!/bin/bash
# round-1 closing captures (the build after the k_intra_rough changes): ncu launch list of the bench command and one
# --set full capture of both intra rough-search kernels; the plain bench line of this build is profiles/r01i_bench_line.json
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
timeout 120 $CMD > $O/plain_r01i.log 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r01i.csv $CMD > $O/ncu_l_r01i.log 2>&1
timeout 60 python scratch/new_kernels.py 3 > $O/new_kernels_r01i.log 2>&1 || exit 1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_intra_rough" --launch-skip 2 --launch-count 2 -f \
    -o $O/prof_r01i_intra python scratch/new_kernels.py 1 > $O/ncu_f_r01i_intra.log 2>&1
tail -3 $O/new_kernels_r01i.log
ls -la $O
