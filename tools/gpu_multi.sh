#!/bin/bash
# multi-GPU: NCCL sharding tests + strong-scaling bench lines at N = 4, 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q ) > gpurun_out/pytest_mgpu.log 2>&1
tail -4 gpurun_out/pytest_mgpu.log
for n in 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
     bench.py --gpus $n --no-e2e --no-solve > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -1 gpurun_out/bench_n$n.json | cut -c1-400
done
