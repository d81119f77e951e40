#!/bin/bash
mkdir -p gpurun_out
( timeout 420 python -u tools/hex_debug.py ) > gpurun_out/r02_hex_debug.log 2>&1
cat gpurun_out/r02_hex_debug.log
( timeout 300 python -u -m pytest tests/test_host_gpu.py -m gpu -v --timeout 120 -p no:cacheprovider -x -k "default_face or temperature" ) > gpurun_out/r02_pytest_host.log 2>&1
tail -40 gpurun_out/r02_pytest_host.log
( timeout 240 python -u -m pytest tests/test_parity_gpu.py -m gpu -v --timeout 100 -p no:cacheprovider --durations=0 -k "hex_core or delta_slabs or delta_golden or anderson" ) > gpurun_out/r02_pytest_sel.log 2>&1
tail -60 gpurun_out/r02_pytest_sel.log
