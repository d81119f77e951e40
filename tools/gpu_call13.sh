#!/bin/bash
for dbg in 128 192 256 384; do
  PAMPA_SN_DBG=$dbg timeout 120 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/bench_t.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t.json").read().strip().splitlines()[-1])
    print("pub=%d"%($dbg//16), "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"])
except Exception as e:
    print("failed", e)
PY
done
# one rank of an 8-way sharded run (gm = 1, short tasks): does the publish interval matter for the critical path?
for dbg in 64 128 256; do
  PAMPA_SN_DBG=$dbg timeout 120 python bench.py --no-cpu-baseline --no-e2e --no-solve --opts '{"rank":0,"num_ranks":8,"shard_mode":1}' > gpurun_out/bench_t.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t.json").read().strip().splitlines()[-1])
    print("8-way rank, pub=%d"%($dbg//16), "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"])
except Exception as e:
    print("failed", e)
PY
done
