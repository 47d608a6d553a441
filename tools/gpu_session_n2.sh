#!/bin/bash
# 2-GPU session: NCCL row-shard parity test, two-device single-process test, bench at N=2 (f32 + bf16, incl. the row-shard record)
T=${1:-r02n2}
timeout 600 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_round2.py::test_two_devices_in_one_process -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_tests.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_f32.json 2> gpurun_out/${T}_bench_f32.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --dtype bf16 > gpurun_out/${T}_bench_bf16.json 2> gpurun_out/${T}_bench_bf16.err
cat gpurun_out/${T}_tests.txt
python - <<PY
import json
for f in ("gpurun_out/${T}_bench_f32.json", "gpurun_out/${T}_bench_bf16.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "value", "%.3e" % d["value"], "e2e", "%.3e" % d["e2e"]["value"], d["e2e"].get("cuda_graph_pipelined"))
        print("    row_shard", d.get("row_shard"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-2500:])
PY
