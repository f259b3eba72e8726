#!/bin/bash
# gprof flat profile of the hooked encoder (build/hm_pg: the same sources compiled with -pg) on the 1080p LDP sequence
set -e
cd "$(dirname "$0")/.."
N=${1:-3}
D=gpurun_out/hm1080pg; mkdir -p $D
python - <<PY
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, synth
seq = synth.make_sequence(1920,1080,$N)
with open('$D/in.yuv','wb') as f:
    for y,u,v in seq:
        f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
PY
ARGS="-c build/hm/cfg/${CFG:-encoder_lowdelay_P_main.cfg} -i $D/in.yuv -wdt 1920 -hgt 1080 -fr 30 -f $N"
( cd $D && env TVC_HM=${TVC_HM:-me,frac,tables,candgrid} LD_LIBRARY_PATH=../../thevc_b200/lib ../../build/hm_pg/TAppEncoderCuda $(echo $ARGS | sed "s# build/# ../../build/#; s#$D/#./#g") -b out.bin > enc.log 2> enc.err )
gprof -b -p build/hm_pg/TAppEncoderCuda $D/gmon.out 2>/dev/null | head -45 > $D/gprof_flat.txt
cat $D/gprof_flat.txt
rm -f $D/in.yuv $D/out.bin $D/gmon.out
