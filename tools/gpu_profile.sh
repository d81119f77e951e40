#!/bin/bash
# round profile: bench line, ncu launch list and one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sn_sweep_flow -s 1 -c 1 -o gpurun_out/r01_flow_full -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/flow_full.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:sn_unshear -s 1 -c 1 -o gpurun_out/r01_unshear_full -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/unshear_full.log 2>&1
cat gpurun_out/r01_bench_n1.json
