#!/bin/bash
# correctness (small case) and C4 timing of the publish-before-store variants; dbg: pub<<4 | flags
timeout 300 python tools/cmp_gm.py 64 64 216 '[{"wave_launch":1},{"group_merge":8,"dbg":16},{"group_merge":8,"dbg":16},{"group_merge":8},{"group_merge":8},{"group_merge":4,"dbg":24},{"group_merge":8,"dbg":24},{"group_merge":8,"dbg":64},{"group_merge":1,"dbg":16}]'
for dbg in 0 16 64 128 256 24 40 72 2; do
  PAMPA_SN_DBG=$dbg timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/bench_dbg$dbg.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_dbg$dbg.json").read().strip().splitlines()[-1])
    print("dbg=$dbg", "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"], "k", d["config"]["keff_after_steps"])
except Exception as e:
    print("dbg=$dbg failed", e)
PY
done
