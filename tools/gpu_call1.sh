#!/bin/bash
# one gpurun call: GPU tests, bench (default + A/B), ncu launch list, ncu full capture of the sweep kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve --opts '{"wave_launch":1}' > gpurun_out/bench_wave.json 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve --opts '{"store_psi":0}' > gpurun_out/bench_nopsi.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sn_sweep_flow -s 1 -c 1 -o gpurun_out/flow_full -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/flow_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_n1.json
