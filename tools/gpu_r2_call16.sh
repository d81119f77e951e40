#!/bin/bash
# round 2: full GPU suite, bench line (fused tail and separate passes), launch list with DRAM bytes per kernel,
# full captures of the flow kernel and of the fused un-shear kernel
mkdir -p gpurun_out
( timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_gpu.log | head -60
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
PAMPA_SN_NO_FUSE=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02_bench_n1_nofuse.json 2> gpurun_out/r02_bench_n1_nofuse.err
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02_bench_hex_s8.json 2> gpurun_out/r02_bench_hex_s8.err
for f in r02_bench_n1 r02_bench_n1_nofuse r02_bench_hex_s8; do
python - $f <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g ms/step %.2f frac %.3f phases %s e2e %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_phases_ms"], d.get("e2e") and d["e2e"]["value"], d.get("keff_solve")))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/%s.err" % sys.argv[1]).read()[-2000:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r02_launches_hex.csv python bench.py --mesh hex --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches_hex.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sn_sweep_flow -s 1 -c 1 -o gpurun_out/r02_flow_full -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/flow_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sn_unshear -s 1 -c 1 -o gpurun_out/r02_unshear_full -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/unshear_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sn_sweep_flow -s 2 -c 1 -o gpurun_out/r02_flow_hex_full -f \
   python bench.py --mesh hex --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/flow_hex_full.log 2>&1
for r in r02_flow_full r02_unshear_full r02_flow_hex_full; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/$r.details.csv 2>/dev/null
done
ncu -i gpurun_out/r02_unshear_full.ncu-rep --page source --csv > gpurun_out/r02_unshear_full.source.csv 2>/dev/null
ncu -i gpurun_out/r02_flow_hex_full.ncu-rep --page source --csv > gpurun_out/r02_flow_hex_full.source.csv 2>/dev/null
rm -f gpurun_out/r02_unshear_full.ncu-rep gpurun_out/r02_flow_hex_full.ncu-rep
ls -la gpurun_out/ | grep r02_ | tail -30
