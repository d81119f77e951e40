#!/bin/bash
# round 2: un-shear kernel templated on the z split (the one-slab variant is the round's earlier kernel again):
# parity subset, bench line of record, launch list
mkdir -p gpurun_out
( timeout 600 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "fused or hex_schedule or anderson or reduced or keff_hex" ) > gpurun_out/r02_pytest_sel.log 2>&1
tail -3 gpurun_out/r02_pytest_sel.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_sel.log | head -80
timeout 600 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_n1.json").read().strip().splitlines()[-1])
    print("bench value %.4g ms/step %.2f frac %.3f phases %s e2e %s solve %s cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["step_phases_ms"].items()}, d["e2e"]["value"], {k: d["keff_solve"][k] for k in ("wall_s", "iterations", "keff", "ms_per_iteration")}, d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02f_bench_n1.err").read()[-2000:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
PAMPA_SN_UNSHEAR_ZSPLIT=2 timeout 200 python bench.py --groups 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02f_g1_z2.json 2> gpurun_out/r02f_g1_z2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f_g1_z2.json").read().strip().splitlines()[-1]); p = d["step_phases_ms"]
print("one group, two slabs: ms/step %.3f kernel %.3f layout passes %.3f" % (d["ms_per_step"], p["sweep kernel alone"], p["shear + sweep + un-shear"] - p["sweep kernel alone"]))
PY
