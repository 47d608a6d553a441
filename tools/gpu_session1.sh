#!/bin/bash
# one GPU session: tests, smoke, bench lines (+ reference arm, Mlp-max VarMisuse workload), launch list, full ncu captures of the dominant kernels
T=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py::test_two_devices_in_one_process 2>&1 | tail -60 > gpurun_out/${T}_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_f32.json 2> gpurun_out/${T}_bench_f32.err
timeout 300 python bench.py --steps 20 --warmup 5 --dtype bf16 > gpurun_out/${T}_bench_bf16.json 2> gpurun_out/${T}_bench_bf16.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 --workload varmisuse --agg max --layers mlp > gpurun_out/${T}_bench_varmisuse_max.json 2> gpurun_out/${T}_bench_varmisuse_max.err
timeout 200 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.txt 2>&1
PTGNN_FUSED_TRACE=1 python tools/fused_trace.py f32 > gpurun_out/${T}_trace_f32.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_aggregate -s 10 -c 2 -o gpurun_out/${T}_fused_f32 python bench.py --profile --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_aggregate -s 10 -c 2 -o gpurun_out/${T}_fused_bf16 python bench.py --profile --steps 2 --warmup 3 --dtype bf16 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^gru_ws -s 10 -c 1 -o gpurun_out/${T}_gru_f32 python bench.py --profile --steps 2 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^gru_ws -s 10 -c 1 -o gpurun_out/${T}_gru_bf16 python bench.py --profile --steps 2 --warmup 3 --dtype bf16 > /dev/null 2>&1
cat gpurun_out/${T}_tests.txt | tail -15
tail -3 gpurun_out/${T}_smoke.txt
python - <<PY
import json
for f in ("gpurun_out/${T}_bench_f32.json", "gpurun_out/${T}_bench_bf16.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 3), "value", "%.3e" % d["value"], "e2e", "%.3e" % d["e2e"]["value"], "serial ms", round(d["e2e"]["serial_ms_per_step"], 2),
              "host ms", round(d["e2e"]["host_enqueue_ms_per_step"], 2), "launches/step", d["gpu_launches_per_step"])
        for k, v in d["kernels"].items():
            print("   ", k, round(v["avg_ms"], 4), "x", v["launches_per_step"], "share", round(v["share_of_step"], 3), "frac_hbm", round(v.get("frac_hbm", 0), 3), "frac_tensor", round(v.get("frac_tensor_exact_peak", 0), 3))
        print("    roofline", d["roofline"]["bound"], round(d["roofline"]["frac"], 3), "layer frac", round(d["layer_roofline"]["frac"], 4), "cpu", d.get("cpu_baseline") and "%.3e" % d["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
