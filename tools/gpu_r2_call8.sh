#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --size 96 96 96 --steps 3 --no-cpu-baseline > gpurun_out/r02_smoke_a.json 2> gpurun_out/r02_smoke_a.err
tail -c 1500 gpurun_out/r02_smoke_a.json; tail -5 gpurun_out/r02_smoke_a.err
timeout 200 python bench.py --mesh hex --rings 60 --size 1 1 40 --order 12 --groups 16 --opts '{"store_psi":0}' --no-solve --steps 3 > gpurun_out/r02_smoke_b.json 2> gpurun_out/r02_smoke_b.err
tail -c 1500 gpurun_out/r02_smoke_b.json; tail -5 gpurun_out/r02_smoke_b.err
