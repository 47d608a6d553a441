#!/bin/bash
# A/B: previous library build vs the current one (fused kernel time + whole step), then the fused parity tests
T=${1:-ab3}
OUT=gpurun_out/${T}_ab3.txt
: > $OUT
for d in f32 bf16; do
  for rep in 1 2; do
    PTGNN_TOOLS_LIB=tools/_variants/libptgnn_b200_prev.so python tools/fused_time.py $d prev >> $OUT 2>&1
    python tools/fused_time.py $d new >> $OUT 2>&1
  done
  PTGNN_TOOLS_LIB=tools/_variants/libptgnn_b200_prev.so python tools/step_time.py $d prev >> $OUT 2>&1
  python tools/step_time.py $d new >> $OUT 2>&1
done
cat $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_round2.py -m gpu -q -x --deselect tests/test_gpu_round2.py::test_two_devices_in_one_process 2>&1 | tail -4
