#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-solve > gpurun_out/bench_a.json 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve --opts '{"dt_max":10}' > gpurun_out/bench_dt10.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches2.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_a.json gpurun_out/bench_dt10.json
grep -o '"sn_[a-z_]*\|void sn_[a-z_<0-9, >]*\|"ns","[0-9]*"' gpurun_out/launches2.csv | paste - - | tail -12
