#!/bin/bash
# round 2: z-split of the un-shear passes: parity subset, and the one-group problem a rank of an 8-GPU run sees
mkdir -p gpurun_out
( timeout 600 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "fused or hex or anderson or schedule or reduced" ) > gpurun_out/r02_pytest_sel.log 2>&1
tail -3 gpurun_out/r02_pytest_sel.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_sel.log | head -80
show() {
python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    p = d["step_phases_ms"]
    print(sys.argv[1], "ms/step %.3f kernel %.3f layout passes %.3f source %.3f reduce %.3f keff %s" % (d["ms_per_step"], p["sweep kernel alone"], p["shear + sweep + un-shear"] - p["sweep kernel alone"], p["source"], p["reduce + scalars + exchange"], d["config"]["keff_after_steps"]))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/%s.err" % sys.argv[1]).read()[-2000:])
PY
}
for z in 1 2 3 4; do
PAMPA_SN_UNSHEAR_ZSPLIT=$z timeout 200 python bench.py --groups 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02c_g1_z$z.json 2> gpurun_out/r02c_g1_z$z.err; show r02c_g1_z$z
done
timeout 200 python bench.py --groups 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02c_g1_auto.json 2> gpurun_out/r02c_g1_auto.err; show r02c_g1_auto
PAMPA_SN_NO_FUSE=1 timeout 200 python bench.py --groups 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02c_g1_nofuse.json 2> gpurun_out/r02c_g1_nofuse.err; show r02c_g1_nofuse
for z in 1 2; do
PAMPA_SN_UNSHEAR_ZSPLIT=$z timeout 200 python bench.py --groups 2 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02c_g2_z$z.json 2> gpurun_out/r02c_g2_z$z.err; show r02c_g2_z$z
done
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02c_n1.json 2> gpurun_out/r02c_n1.err; show r02c_n1
