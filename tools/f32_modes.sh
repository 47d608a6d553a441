#!/bin/bash
# fp32 bench in both tcgen05 operand modes (TS = A from tensor memory, SS = A hi/lo from shared memory)
for m in ${MODES:-ts ss}; do
  echo "== mode $m"
  PTGNN_TC_MODE=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['kernels']
print('ms/step %.3f  msg %.4f reduce %.4f gru %.4f' % (j['ms_per_step'], k['message']['avg_ms'], k['reduce']['avg_ms'], k['gru']['avg_ms']))"
done
