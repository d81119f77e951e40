#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -u -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 240 -p no:cacheprovider -k "schedule or fused or full_size or anderson" ) > gpurun_out/r02_pytest_sel.log 2>&1
tail -5 gpurun_out/r02_pytest_sel.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_sel.log | head -60
for f in "" "PAMPA_SN_NO_FUSE=1"; do
env $f timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02_bench_tmp.json 2> gpurun_out/r02_bench_tmp.err
python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02_bench_tmp.json").read().strip().splitlines()[-1])
    print("C4", sys.argv[1] or "fused", "value %.4g ms/step %.2f phases %s keff %s" % (d["value"], d["ms_per_step"], d["step_phases_ms"], d["config"]["keff_after_steps"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench_tmp.err").read()[-2000:])
PY
done
