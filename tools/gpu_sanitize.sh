#!/bin/bash
# memcheck (and an informational racecheck) of the sweep kernels on a small Cartesian core
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/cmp_gm.py 48 40 24 '[{"wave_launch":1},{},{"group_merge":3},{"store_psi":0},{"inline_edges":1},{"generic_only":1}]' > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/cmp_gm.py 32 32 12 '[{"wave_launch":1},{}]' > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/racecheck.log; tail -6 gpurun_out/racecheck.log
