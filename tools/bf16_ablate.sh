#!/bin/bash
# ablation of the bf16 pipeline: which role bounds each kernel (PTGNN_TC_DEBUG bits: 1 no MMA, 2 no store, 4 no loads, 8 no drain)
for d in ${ABL:-0 1 2 4 10 5 7 11 14 13}; do
  echo "== dbg $d"
  PTGNN_TC_DEBUG=$d timeout 120 python bench.py --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['kernels']
print('msg %.4f reduce %.4f gru %.4f' % (k['message']['avg_ms'], k['reduce']['avg_ms'], k['gru']['avg_ms']))"
done
