#!/bin/bash
# quick session: fused parity tests, timeline trace, bench lines
T=${1:-r02g}
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_round2.py tests/test_gpu_bf16.py tests/test_gpu_layers.py -m gpu -q -x --deselect tests/test_gpu_round2.py::test_two_devices_in_one_process 2>&1 | tail -8 > gpurun_out/${T}_tests.txt
python tools/fused_trace.py f32 > gpurun_out/${T}_trace_f32.txt 2>&1
python tools/fused_trace.py bf16 > gpurun_out/${T}_trace_bf16.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_f32.json 2> gpurun_out/${T}_bench_f32.err
timeout 300 python bench.py --steps 20 --warmup 5 --dtype bf16 > gpurun_out/${T}_bench_bf16.json 2> gpurun_out/${T}_bench_bf16.err
cat gpurun_out/${T}_tests.txt; cat gpurun_out/${T}_trace_f32.txt; grep -A4 "mma (per step)" gpurun_out/${T}_trace_bf16.txt; grep "write-out" gpurun_out/${T}_trace_bf16.txt
python - <<PY
import json
for f in ("gpurun_out/${T}_bench_f32.json", "gpurun_out/${T}_bench_bf16.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 3), "value", "%.3e" % d["value"], "e2e", "%.3e" % d["e2e"]["value"])
        for k, v in d["kernels"].items():
            print("   ", k, round(v["avg_ms"], 4), "x", v["launches_per_step"], "share", round(v["share_of_step"], 3))
        print("    row_shard", d.get("row_shard") and round(d["row_shard"]["ms_per_step"], 3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
