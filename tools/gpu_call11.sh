#!/bin/bash
# one rank of an N-way group-sharded run, emulated on one GPU without the exchange: per-rank compute time
for nr in 8 4; do
  timeout 200 python bench.py --no-cpu-baseline --no-solve --no-e2e --opts "{\"rank\":0,\"num_ranks\":$nr,\"shard_mode\":1}" > gpurun_out/bench_t.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t.json").read().strip().splitlines()[-1])
    print('ranks $nr', "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_t.json").read()[-1500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r8.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve --opts '{"rank":0,"num_ranks":8,"shard_mode":1}' > gpurun_out/launches_r8.log 2>&1
grep -o '"sn_[a-z_]*\|void sn_[a-z_<0-9, >]*\|"ns","[0-9]*"' gpurun_out/launches_r8.csv | paste - - | tail -8
