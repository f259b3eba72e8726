#!/bin/bash
# round 2, call N: CU-level first search in the group kernel, look-ahead census groups in the encoder shim
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02n_parity.log 2>&1; echo "parity rc=$?"; tail -6 $O/r02n_parity.log
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu > $O/r02n_props.log 2>&1; echo "props rc=$?"; tail -3 $O/r02n_props.log
( timeout 1200 python -m pytest tests/test_hm_md5.py -x -q -m gpu -k "census_lookup or frame_prepass or long_sequences or (test_bitstream_md5_identical_to_reference and lowdelay_P)" -s > $O/r02n_md5.log 2>&1; echo "md5 rc=$?" >> $O/r02n_md5.log ) &
TVC_GROUP_CU=0 timeout 600 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu > $O/r02n_bench_cu0.json 2> $O/r02n_bench_cu0.err; echo "bench cu0 rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02n_bench.json 2> $O/r02n_bench.err; echo "bench rc=$?"; tail -3 $O/r02n_bench.err
python - <<'P'
import json
for f in ("gpurun_out/r02n_bench_cu0.json", "gpurun_out/r02n_bench.json"):
    b = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, b["ms_per_step"], b["detail"]["phase_ms_per_step"]["me_search"], b["e2e"]["ms_per_step"])
    h = b["detail"].get("hm_encode")
    if h: print({k: h[k] for k in h if k != "hooks"}); print("\n".join(h["hooks"]))
P
wait
tail -12 $O/r02n_md5.log
