#!/bin/bash
# round 2, call L: SAD routine variants of the group kernel (4-chain generic vs width-specialised), MINB 2/3; PSNR hook test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02l_parity.log 2>&1; echo "parity rc=$?"; tail -3 $O/r02l_parity.log
for v in "0 2" "1 2" "0 3" "1 3"; do set -- $v
  TVC_GROUP_SAD=$1 TVC_GROUP_MINB=$2 timeout 600 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu > $O/r02l_bench_sad$1_minb$2.json 2> $O/r02l_bench_sad$1_minb$2.err; echo "bench sad=$1 minb=$2 rc=$?"
  python -c "
import json; b=json.loads(open('$O/r02l_bench_sad$1_minb$2.json').read().strip().splitlines()[-1]); print('sad=$1 minb=$2', b['ms_per_step'], b['detail']['phase_ms_per_step']['me_search'])"
done
TVC_GROUP_SAD=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_frame_prepass or me_group_search" > $O/r02l_parity_sad1.log 2>&1; echo "parity sad1 rc=$?"; tail -3 $O/r02l_parity_sad1.log
timeout 600 python -m pytest tests/test_hm_md5.py -x -q -m gpu -k "psnr" -s > $O/r02l_psnr.log 2>&1; echo "psnr rc=$?"; tail -4 $O/r02l_psnr.log
