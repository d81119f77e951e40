#!/bin/bash
# correctness (small case, publish every step = dbg 16) and C4 timing of the publish variants
timeout 500 python tools/cmp_gm.py 64 64 216 '[{"wave_launch":1},{"group_merge":8,"dbg":24},{"group_merge":8,"dbg":24},{"group_merge":8,"dbg":4112},{"group_merge":8,"dbg":4112},{"group_merge":8,"dbg":4120},{"group_merge":8,"dbg":18},{"group_merge":8,"dbg":40},{"group_merge":8,"dbg":4128}]'
for dbg in 0 8 4096 2 66 130 258 4160 4224; do
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
