#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -u -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider -x ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_gpu.log | head -80
timeout 300 python bench.py > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1_d.json").read().strip().splitlines()[-1])
    print("C4 value %.4g ms/step %.2f kernel frac %.3f phases %s e2e %.4g solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_phases_ms"], d["e2e"]["value"], d["keff_solve"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench_n1_d.err").read()[-2000:])
PY
PAMPA_SN_NO_FUSE=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02_bench_n1_nofuse.json 2> gpurun_out/r02_bench_n1_nofuse.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1_nofuse.json").read().strip().splitlines()[-1])
    print("C4 no-fuse value %.4g ms/step %.2f phases %s keff %s" % (d["value"], d["ms_per_step"], d["step_phases_ms"], d["config"]["keff_after_steps"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench_n1_nofuse.err").read()[-2000:])
PY
