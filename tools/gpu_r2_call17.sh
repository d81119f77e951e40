#!/bin/bash
# round 2: fused tail on every tiling / in-place fused reduction of the accelerated iteration: full GPU suite, bench
# lines (Cartesian with the k-eff solve, hexagonal), launch list of the accelerated iteration, Anderson depth scan
mkdir -p gpurun_out
( timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_gpu.log | head -80
show() {
python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g ms/step %.2f frac %.3f phases %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["step_phases_ms"].items()}, d.get("keff_solve") and {k: d["keff_solve"][k] for k in ("wall_s", "iterations", "keff", "ms_per_iteration")}))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/%s.err" % sys.argv[1]).read()[-2000:])
PY
}
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; show r02b_bench_n1
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_hex_s8.json 2> gpurun_out/r02b_bench_hex_s8.err; show r02b_bench_hex_s8
for dpt in 3 5; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --opts "{\"anderson_depth\": $dpt}" > gpurun_out/r02b_bench_aa$dpt.json 2> gpurun_out/r02b_bench_aa$dpt.err; show r02b_bench_aa$dpt
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv \
   --log-file gpurun_out/r02_launches_solve.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_solve.log 2>&1
tail -2 gpurun_out/launches_solve.log
