#!/bin/bash
( timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_host_gpu.py -m gpu -x -q -k "not full_size and not schedule" ) > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
timeout 100 python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['keff_solve'])"
