#!/bin/bash
# round 2, 8 GPUs, final: C4 at N = 8 / 4 / 2 and config 5 at N = 8 with the peer-to-peer exchange
mkdir -p gpurun_out
( timeout 200 python -u -m pytest tests/test_sharding.py -m gpu -q --timeout 150 -p no:cacheprovider ) > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -3 gpurun_out/r02_pytest_mgpu.log
runn() {
  n=$1; tag=$2; to=$3; shift 3
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
     bench.py --gpus $n "$@" > gpurun_out/r02_n${n}_$tag.json 2> gpurun_out/r02_n${n}_$tag.err
  python - "$n" "$tag" <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/r02_n%s_%s.json" % (n, tag)).read().strip().splitlines()[-1])
    print("N=%s" % n, tag, "| value %.4g ms/step %.2f kernel frac %.3f phases %s e2e %s parity ok %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["step_phases_ms"], d["e2e"] and ("%.4g" % d["e2e"]["value"], d["e2e"]["phases_ms"]), d["sharded_parity"]["ok"], d["keff_solve"]))
except Exception as e:
    print("N=%s" % n, tag, "failed", e); print("\n".join(l for l in open("gpurun_out/r02_n%s_%s.err" % (n, tag)).read().splitlines() if "rank" in l or "Error" in l)[-3000:])
PY
}
runn 8 c4_final 300
runn 8 c5_final 600 --config c5 --steps 5
runn 4 c4_final 240
runn 2 c4_final 240
