#!/bin/bash
# backward parity + tcgen05 peak micro-benchmark
T=${1:-r02o}
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/${T}_tests.txt
timeout 300 python tools/tc_peak.py > gpurun_out/${T}_tc_peak.txt 2>&1
cat gpurun_out/${T}_tests.txt gpurun_out/${T}_tc_peak.txt
