#!/bin/bash
# round 2, 2 GPUs, final code: NCCL / peer-to-peer parity tests
mkdir -p gpurun_out
( timeout 400 python -u -m pytest tests/test_sharding.py -m gpu -x -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -3 gpurun_out/r02_pytest_mgpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_mgpu.log | head -60
