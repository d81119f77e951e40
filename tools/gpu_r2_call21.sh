#!/bin/bash
# round 2, final kernels: full GPU suite, the bench line of record, ncu launch list, compute-sanitizer on the new paths
mkdir -p gpurun_out
( timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_gpu.log | head -80
timeout 600 python bench.py > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02e_bench_n1.json").read().strip().splitlines()[-1])
    print("bench value %.4g ms/step %.2f frac %.3f phases %s e2e %s solve %s cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["step_phases_ms"].items()}, d["e2e"]["value"], {k: d["keff_solve"][k] for k in ("wall_s", "iterations", "keff", "ms_per_iteration")}, d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02e_bench_n1.err").read()[-2000:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r02e_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_r2.py > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -14 gpurun_out/r02_memcheck.log
