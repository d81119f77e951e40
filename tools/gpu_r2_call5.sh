#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -u tools/refl_debug.py ) > gpurun_out/r02_refl_debug.log 2>&1
cat gpurun_out/r02_refl_debug.log
( timeout 200 python -u -m pytest tests/test_host_gpu.py -m gpu -v --timeout 100 -p no:cacheprovider -k "partitioned or feedback" ) > gpurun_out/r02_pytest_part.log 2>&1
grep -E "PASSED|FAILED|ERROR|passed|failed" gpurun_out/r02_pytest_part.log | tail; grep -B2 -A30 "^E " gpurun_out/r02_pytest_part.log | head -60
timeout 300 python bench.py > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1_c.json").read().strip().splitlines()[-1])
    print("C4 value %.4g ms/step %.2f kernel frac %.3f sweep %.2f e2e %.4g solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["sweep_ms_per_step"], d["e2e"]["value"], d["keff_solve"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench_n1_c.err").read()[-2000:])
PY
timeout 200 python bench.py --mesh hex --no-e2e --steps 5 --order 8 --rings 120 --size 1 1 100 > gpurun_out/r02_hex_s8_sorted.json 2> gpurun_out/r02_hex_s8_sorted.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_hex_s8_sorted.json").read().strip().splitlines()[-1])
    print("hex | value %.4g ms/step %.2f kernel ms %.2f frac %.3f with layout %.3f solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"] * d["roofline"]["launches_per_step"], d["roofline"]["frac"], d["roofline"]["frac_with_layout_passes"], d["keff_solve"]))
except Exception as e:
    print("hex failed", e); print(open("gpurun_out/r02_hex_s8_sorted.err").read()[-1500:])
PY
