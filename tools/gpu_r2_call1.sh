#!/bin/bash
# round 2, call 1: full GPU test-suite, C4 bench line, hex bench lines (dataflow kernel on lattice tilings), memcheck
mkdir -p gpurun_out
nvidia-smi -L
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -25 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1_a.json").read().strip().splitlines()[-1])
    print("C4 value %.4g ms/step %.2f kernel frac %.3f sweep %.2f e2e %.4g solve %s cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["sweep_ms_per_step"], d["e2e"]["value"], d["keff_solve"], d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_bench_n1_a.err").read()[-2000:])
PY
for o in '{}' '{"dt_max":3}' '{"dt_max":5}' '{"store_psi":0}' '{"generic_only":1}'; do
  tag=$(echo "$o" | tr -d '{}":' | tr ',' '_'); [ -z "$tag" ] && tag=default
  timeout 400 python bench.py --mesh hex --order 8 --rings 120 --size 1 1 100 --no-e2e --steps 5 --opts "$o" > gpurun_out/r02_hex_$tag.json 2> gpurun_out/r02_hex_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02_hex_%s.json" % tag).read().strip().splitlines()[-1])
    print("hex", tag, "| value %.4g ms/step %.2f kernel ms %.2f frac %.3f with layout %.3f launches %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"] * d["roofline"]["launches_per_step"], d["roofline"]["frac"], d["roofline"]["frac_with_layout_passes"], d["roofline"]["launches_per_step"], d["keff_solve"]))
except Exception as e:
    print("hex", tag, "failed", e); print(open("gpurun_out/r02_hex_%s.err" % tag).read()[-1500:])
PY
done
timeout 300 python bench.py --mesh hex --order 12 --rings 80 --size 1 1 100 --groups 16 --no-e2e --steps 5 > gpurun_out/r02_hex_s12.json 2> gpurun_out/r02_hex_s12.err
tail -c 600 gpurun_out/r02_hex_s12.json
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/hex_check.py > gpurun_out/r02_memcheck_hex.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/r02_memcheck_hex.log
