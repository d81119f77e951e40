#!/bin/bash
# round 2, final code: ncu launch lists (C4 and the hexagonal lattice) with DRAM bytes per kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv \
   --log-file gpurun_out/r02p_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv \
   --log-file gpurun_out/r02p_launches_hex.csv python bench.py --mesh hex --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches_hex.log 2>&1
ls -la gpurun_out/r02p_*
