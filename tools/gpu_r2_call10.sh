#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -u -m pytest tests/test_sharding.py -m gpu -v --timeout 200 -p no:cacheprovider ) > gpurun_out/r02_pytest_mgpu.log 2>&1
grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed" gpurun_out/r02_pytest_mgpu.log | tail; grep -B5 -A25 "Error\|assert" gpurun_out/r02_pytest_mgpu.log | head -80
run2() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 "$@" > gpurun_out/r02_n2_$tag.json 2> gpurun_out/r02_n2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02_n2_%s.json" % tag).read().strip().splitlines()[-1])
    print(tag, "| value %.4g ms/step %.2f frac %.3f phases %s e2e %s parity ok %s keff after steps %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_phases_ms"], d["e2e"] and ("%.4g" % d["e2e"]["value"]), d["sharded_parity"]["ok"], d["config"]["keff_after_steps"], d["keff_solve"]))
except Exception as e:
    print(tag, "failed", e); print("\n".join(l for l in open("gpurun_out/r02_n2_%s.err" % tag).read().splitlines() if "rank" in l or "Error" in l or "pampa" in l)[-2500:])
PY
}
run2 c4_p2p --steps 10 --opts '{"verbose":1}'
PAMPA_SN_NO_P2P=1 run2 c4_nccl --steps 10
