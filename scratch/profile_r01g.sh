#!/bin/bash
# round-1 final captures: plain bench line, ncu launch list of the same command, one --set full capture per hot kernel
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0"
python bench.py > $O/bench_r01g.json 2> $O/bench_r01g.err || exit 1
$CMD > $O/plain_r01g.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r01g.csv $CMD > $O/ncu_l_r01g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_me_sad_tables|k_me_raster|k_me_search|k_me_frac|k_rdoq" \
    --launch-skip 9 --launch-count 9 -f -o $O/prof_r01g $CMD > $O/ncu_f_r01g.log 2>&1
python scratch/new_kernels.py 3 > $O/new_kernels_r01g.log 2>&1
for k in k_ctu_cost_grids k_pred_cost k_intra_rough; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 --launch-count 1 -f -o $O/prof_r01g_$k python scratch/new_kernels.py 1 > $O/ncu_f_r01g_$k.log 2>&1
done
tail -3 $O/new_kernels_r01g.log
python -c "
import json; d=json.load(open('$O/bench_r01g.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['cpu_baseline'] and d['cpu_baseline']['value'], d['detail'].get('hm_encode'), d['detail'].get('intra_rough'))"
