#!/bin/bash
# round 2, 2 GPUs: NCCL / peer-to-peer parity tests (Cartesian both modes, hexagonal lattice group mode) and the C4 line
mkdir -p gpurun_out
( timeout 400 python -u -m pytest tests/test_sharding.py -m gpu -x -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -3 gpurun_out/r02_pytest_mgpu.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_mgpu.log | head -60
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02b_n2_c4.json 2> gpurun_out/r02b_n2_c4.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02b_n2_c4.json").read().strip().splitlines()[-1])
    print("N=2 | value %.4g ms/step %.2f kernel frac %.3f phases %s e2e %s parity %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_phases_ms"], d["e2e"] and ("%.4g" % d["e2e"]["value"]), d["sharded_parity"], d["keff_solve"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02b_n2_c4.err").read()[-3000:])
PY
