#!/bin/bash
# round 2, final code: full GPU suite, smoke(), the bench line of record, one config-5-shaped line on one GPU
mkdir -p gpurun_out
( timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_gpu.log | head -80
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02o_bench_n1.json").read().strip().splitlines()[-1])
    print("bench value %.4g ms/step %.2f frac %.3f phases %s e2e %s solve %s cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["step_phases_ms"].items()}, d["e2e"]["value"], {k: d["keff_solve"][k] for k in ("wall_s", "iterations", "keff", "ms_per_iteration")}, d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02o_bench_n1.err").read()[-2000:])
PY
timeout 300 python bench.py --mesh hex --rings 80 --size 1 1 100 --order 12 --groups 16 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02o_hex_s12_16g.json 2> gpurun_out/r02o_hex_s12_16g.err
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02o_hex_s8.json 2> gpurun_out/r02o_hex_s8.err
for f in r02o_hex_s12_16g r02o_hex_s8; do
python - $f <<'PY'
import json, sys
d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1]); p = d["step_phases_ms"]
print(sys.argv[1], d["config"]["workload"], "| ms/step %.3f value %.4g kernel %.3f (frac %.3f, with layout passes %.3f) launches/step %s" % (d["ms_per_step"], d["value"], p["sweep kernel alone"], d["roofline"]["frac"], d["roofline"]["frac_with_layout_passes"], d["roofline"]["launches_per_step"]))
PY
done
