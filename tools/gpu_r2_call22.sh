#!/bin/bash
# round 2: compute-sanitizer memcheck on the round-2 code paths (tools/sanitize_r2.py), racecheck on the hexagonal flow kernel
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_r2.py > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 gpurun_out/r02_memcheck.log
