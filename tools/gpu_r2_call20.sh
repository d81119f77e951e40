#!/bin/bash
# round 2, 8 GPUs, final kernels: C4 at N = 8 / 4 and config 5 at N = 8
mkdir -p gpurun_out
runn() {
  n=$1; tag=$2; to=$3; shift 3
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
     bench.py --gpus $n "$@" > gpurun_out/r02d_n${n}_$tag.json 2> gpurun_out/r02d_n${n}_$tag.err
  python - "$n" "$tag" <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/r02d_n%s_%s.json" % (n, tag)).read().strip().splitlines()[-1])
    print("N=%s" % n, tag, "| value %.4g ms/step %.2f kernel frac %.3f phases %s e2e %s parity ok %s solve %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: round(v, 3) for k, v in d["step_phases_ms"].items()}, d["e2e"] and ("%.4g" % d["e2e"]["value"], {k: round(v, 1) for k, v in d["e2e"]["phases_ms"].items()}), d["sharded_parity"]["ok"], d["keff_solve"] and {k: d["keff_solve"][k] for k in ("wall_s", "iterations", "keff", "ms_per_iteration")}))
except Exception as e:
    print("N=%s" % n, tag, "failed", e); print("\n".join(l for l in open("gpurun_out/r02d_n%s_%s.err" % (n, tag)).read().splitlines() if "rank" in l or "Error" in l)[-3000:])
PY
}
runn 8 c4 300 --no-cpu-baseline
runn 8 c5 600 --config c5 --steps 5 --no-cpu-baseline
runn 4 c4 240 --no-cpu-baseline
