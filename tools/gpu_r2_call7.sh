#!/bin/bash
mkdir -p gpurun_out
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
    print(tag, "failed", e); print("\n".join(l for l in open("gpurun_out/r02_n2_%s.err" % tag).read().splitlines() if "rank" in l or "Error" in l)[-2500:])
PY
}
run2 c4_small --size 96 96 96 --steps 5
run2 hex --mesh hex --order 8 --rings 60 --size 1 1 64 --steps 5
run2 c4 --steps 10
