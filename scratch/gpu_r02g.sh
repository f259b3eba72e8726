#!/bin/bash
# round 2, call G: gprof flat profile of the hooked encoder (build/hm_pg = the same sources with -pg), 1 I + 4 P pictures at 1080p, fast hook set
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TVC_HM=me,frac,tables,frame,candgrid timeout 170 bash scratch/hm1080_gprof.sh 5 > gpurun_out/r02g_gprof.log 2>&1; echo "gprof rc=$?"
tail -5 gpurun_out/hm1080pg/enc.log; grep TLibCuda gpurun_out/hm1080pg/enc.err | tail -8
