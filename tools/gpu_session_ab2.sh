#!/bin/bash
T=${1:-ab2}
OUT=gpurun_out/${T}_ab2.txt
: > $OUT
for d in f32 bf16; do
  PTGNN_TOOLS_LIB=tools/_variants/libptgnn_b200_r02e.so python tools/step_time.py $d r02e >> $OUT 2>&1
  python tools/step_time.py $d head >> $OUT 2>&1
  PTGNN_B200_CHAIN=0 python tools/step_time.py $d head >> $OUT 2>&1
  PTGNN_TOOLS_LIB=tools/_variants/libptgnn_b200_r02e.so python tools/step_time.py $d r02e >> $OUT 2>&1
  python tools/step_time.py $d head >> $OUT 2>&1
done
cat $OUT
