#!/bin/bash
# round 2, call 4 (2 GPUs): NCCL sharding tests, sharded bench lines (parity check, partitioned fields, device-side
# Anderson), a sharded hex line
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( timeout 300 python -u -m pytest tests/test_sharding.py -m gpu -v --timeout 200 -p no:cacheprovider ) > gpurun_out/r02_pytest_mgpu.log 2>&1
grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed" gpurun_out/r02_pytest_mgpu.log | tail; grep -B5 -A25 "Error\|assert" gpurun_out/r02_pytest_mgpu.log | head -60
run2() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 "$@" > gpurun_out/r02_n2_$tag.json 2> gpurun_out/r02_n2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02_n2_%s.json" % tag).read().strip().splitlines()[-1])
    print(tag, "| value %.4g ms/step %.2f frac %.3f e2e %s parity %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"] and ("%.4g" % d["e2e"]["value"], d["e2e"]["phases_ms"]), d["sharded_parity"], d["keff_solve"]))
except Exception as e:
    print(tag, "failed", e); print(open("gpurun_out/r02_n2_%s.err" % tag).read()[-2500:])
PY
}
run2 c4_small --size 96 96 96 --steps 5
run2 c4 --steps 10
run2 hex --mesh hex --order 8 --rings 60 --size 1 1 64 --steps 5
