#!/bin/bash
# round 2: racecheck (shared memory) of the three-face dataflow kernel and the layout passes on a small hexagonal lattice
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_r2.py hex-small > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/r02_racecheck.log; grep "Race reported" gpurun_out/r02_racecheck.log | sed 's/.*between//' | sort | uniq -c | sort -rn | head -20; tail -8 gpurun_out/r02_racecheck.log
