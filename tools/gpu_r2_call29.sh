#!/bin/bash
# round 2: un-shear passes over per-z chunk ranges, host-selected 8x2 / 4x4 kernel variants: parity subset, C4 / hexagonal lines
mkdir -p gpurun_out
( timeout 900 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "fused or hex or schedule or single_sweep or reduced or anderson or reference_cases" ) > gpurun_out/r02_pytest_sel.log 2>&1
tail -3 gpurun_out/r02_pytest_sel.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_sel.log | head -80
show() {
python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    p = d["step_phases_ms"]
    print(sys.argv[1], "ms/step %.3f value %.4g kernel %.3f (frac %.3f, with layout %.3f) layout passes %.3f keff %s" % (d["ms_per_step"], d["value"], p["sweep kernel alone"], d["roofline"]["frac"], d["roofline"]["frac_with_layout_passes"], p["shear + sweep + un-shear"] - p["sweep kernel alone"], d["config"]["keff_after_steps"]))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/%s.err" % sys.argv[1]).read()[-2000:])
PY
}
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02k_c4.json 2> gpurun_out/r02k_c4.err; show r02k_c4
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02k_hex_s8.json 2> gpurun_out/r02k_hex_s8.err; show r02k_hex_s8
timeout 300 python bench.py --mesh hex --rings 80 --size 1 1 100 --order 12 --groups 16 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02k_hex_s12_16g.json 2> gpurun_out/r02k_hex_s12_16g.err; show r02k_hex_s12_16g
PAMPA_SN_NO_FUSE=1 timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02k_hex_s8_nofuse.json 2> gpurun_out/r02k_hex_s8_nofuse.err; show r02k_hex_s8_nofuse
