#!/bin/bash
# N-GPU session (default 4): bench with graph shards + the row-shard record, fp32 and bf16
T=${1:-r02n4}
NG=${2:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 20 --warmup 5 --no-train > gpurun_out/${T}_bench_f32.json 2> gpurun_out/${T}_bench_f32.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $NG --steps 20 --warmup 5 --dtype bf16 > gpurun_out/${T}_bench_bf16.json 2> gpurun_out/${T}_bench_bf16.err
python - <<PY
import json
for f in ("gpurun_out/${T}_bench_f32.json", "gpurun_out/${T}_bench_bf16.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "value", "%.3e" % d["value"], "e2e", "%.3e" % d["e2e"]["value"], "host", d["e2e"].get("host_enqueue_ms_per_step"))
        rs = d.get("row_shard") or {}
        print("    row_shard ms", rs.get("ms_per_step"), "allgather ms", rs.get("allgather_ms_per_step"), "value", rs.get("value"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-2500:])
PY
