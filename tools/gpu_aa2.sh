#!/bin/bash
( timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_host_gpu.py -m gpu -x -q -k "not full_size" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/solve_c4.py 2>&1 | tail -1
