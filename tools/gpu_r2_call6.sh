#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -u tools/refl_debug.py ) > gpurun_out/r02_refl_debug.log 2>&1
grep -v "pampa_sn: it" gpurun_out/r02_refl_debug.log
grep "pampa_sn: it" gpurun_out/r02_refl_debug.log | awk 'NR<=30 || NR%10==0' | head -120
grep "pampa_sn: it" gpurun_out/r02_refl_debug.log | tail -5
