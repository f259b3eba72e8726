#!/bin/bash
# round 2, 8-GPU call: configs[3] (4K he10, 8 frame shards over 8 GPUs) and configs[2] (RA 1080p, 5 closed intra periods of 16 over 5 GPUs),
# then the bench line at N = 8 (independent sequences, one rank per GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r02h_gpus.txt; nproc >> $O/r02h_gpus.txt
TVC_MATRIX_CASES=he10_2160_8,ra_1080_66_idr_ip16 timeout 1500 python -m pytest tests/test_config_matrix.py -q -m gpu -s > $O/r02h_matrix_n8.log 2>&1; echo "matrix rc=$?"
tail -5 $O/r02h_matrix_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02h_bench_n8.json 2> $O/r02h_bench_n8.err; echo "bench8 rc=$?"
tail -c 600 $O/r02h_bench_n8.json
