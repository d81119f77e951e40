#!/bin/bash
# round 2: per-kernel times of the layout passes (ncu launch list) after the un-shear restructure
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file gpurun_out/r02l_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
grep -v "^==" gpurun_out/r02l_launches_bench.csv | grep "gpu__time_duration" | cut -d, -f5,15 | sed 's/(SweepGlobals.*"\(.*\)"$/ \1/' | sort | uniq -c | sort -k2 | head -40
for i in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02l_c4_$i.json 2> gpurun_out/r02l_c4_$i.err
python - $i <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02l_c4_%s.json" % sys.argv[1]).read().strip().splitlines()[-1]); p = d["step_phases_ms"]
print("C4 run", sys.argv[1], "ms/step %.3f kernel %.3f layout passes %.3f" % (d["ms_per_step"], p["sweep kernel alone"], p["shear + sweep + un-shear"] - p["sweep kernel alone"]))
PY
done
