#!/bin/bash
# quick fp32 bench line: ms/step and the per-kernel averages (debug helper; bench.py prints the full JSON line)
timeout 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['kernels']
print('ms/step %.3f  msg %.4f reduce %.4f gru %.4f' % (j['ms_per_step'], k['message']['avg_ms'], k['reduce']['avg_ms'], k['gru']['avg_ms']))"
