#!/bin/bash
# round 2, first GPU call: long-sequence md5 tests, full-size configuration matrix, bench line on the corrected workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt; nproc >> gpurun_out/r02a_gpus.txt
( timeout 1500 python -m pytest tests/test_config_matrix.py -q -m gpu -s > gpurun_out/r02a_matrix.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_matrix.log ) &
MPID=$!
timeout 900 python -m pytest tests/test_hm_md5.py tests/test_shard_encode.py -x -q -m gpu -k "long_sequences or intra_period_shards" > gpurun_out/r02a_long.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_long.log
wait $MPID
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02a_matrix.log gpurun_out/r02a_long.log
