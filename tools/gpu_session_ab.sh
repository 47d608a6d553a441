#!/bin/bash
# A/B timing of fused-kernel variants at config 2 (each line: one process)
T=${1:-ab}
OUT=gpurun_out/${T}_ab.txt
: > $OUT
for d in f32 bf16; do
  for v in r02e v3 vB vM vF3; do
    [ -f tools/_variants/libptgnn_b200_$v.so ] && PTGNN_TOOLS_LIB=tools/_variants/libptgnn_b200_$v.so python tools/fused_time.py $d $v >> $OUT 2>&1
  done
  for b in 0; do
    PTGNN_FUSED_DBG=$b python tools/fused_time.py $d head >> $OUT 2>&1
  done
done
cat $OUT
# parity of the experimental variant
PTGNN_TOOLS_LIB=tools/_variants/libptgnn_b200_vF3.so timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -3
