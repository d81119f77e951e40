#!/bin/bash
# round 2, call 3: full GPU suite (verbose, durations, per-test timeout), C4 and hex bench lines
mkdir -p gpurun_out
( timeout 1100 python -u -m pytest tests -m gpu -v --timeout 240 -p no:cacheprovider --durations=15 ) > gpurun_out/r02_pytest_gpu.log 2>&1
grep -E "PASSED|FAILED|ERROR|SKIPPED|Timeout|passed|failed" gpurun_out/r02_pytest_gpu.log | tail -70
sed -n '/slowest/,/short test summary/p' gpurun_out/r02_pytest_gpu.log | head -24
timeout 400 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1_b.json").read().strip().splitlines()[-1])
    print("C4 value %.4g ms/step %.2f kernel frac %.3f sweep %.2f e2e %.4g solve %s cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["sweep_ms_per_step"], d["e2e"]["value"], d["keff_solve"], d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench_n1_b.err").read()[-2000:])
PY
run_hex() {
  tag=$1; shift
  timeout 200 python bench.py --mesh hex --no-e2e --steps 5 "$@" > gpurun_out/r02_hex_$tag.json 2> gpurun_out/r02_hex_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02_hex_%s.json" % tag).read().strip().splitlines()[-1])
    print("hex", tag, "| value %.4g ms/step %.2f kernel ms %.2f frac %.3f with layout %.3f launches %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"] * d["roofline"]["launches_per_step"], d["roofline"]["frac"], d["roofline"]["frac_with_layout_passes"], d["roofline"]["launches_per_step"], d["keff_solve"]))
except Exception as e:
    print("hex", tag, "failed", e); print(open("gpurun_out/r02_hex_%s.err" % tag).read()[-1500:])
PY
}
run_hex s8_default --order 8 --rings 120 --size 1 1 100
run_hex s8_dt3 --order 8 --rings 120 --size 1 1 100 --opts '{"dt_max":3}'
run_hex s8_dt5 --order 8 --rings 120 --size 1 1 100 --opts '{"dt_max":5}'
run_hex s12_16g --order 12 --rings 80 --size 1 1 100 --groups 16
run_hex s12_16g_nopsi --order 12 --rings 80 --size 1 1 100 --groups 16 --opts '{"store_psi":0}'
run_hex s12_16g_nopsi_dt5 --order 12 --rings 80 --size 1 1 100 --groups 16 --opts '{"store_psi":0,"dt_max":5}'
