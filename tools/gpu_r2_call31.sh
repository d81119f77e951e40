#!/bin/bash
# round 2: un-shear loads chunk-major again: C4 and hexagonal lines
mkdir -p gpurun_out
show() {
python - $1 <<'PY'
import json, sys
d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1]); p = d["step_phases_ms"]
print(sys.argv[1], "ms/step %.3f value %.4g kernel %.3f (frac %.3f, with layout %.3f) layout passes %.3f keff %s" % (d["ms_per_step"], d["value"], p["sweep kernel alone"], d["roofline"]["frac"], d["roofline"]["frac_with_layout_passes"], p["shear + sweep + un-shear"] - p["sweep kernel alone"], d["config"]["keff_after_steps"]))
PY
}
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02m_c4.json 2> gpurun_out/r02m_c4.err; show r02m_c4
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02m_hex_s8.json 2> gpurun_out/r02m_hex_s8.err; show r02m_hex_s8
timeout 300 python bench.py --mesh hex --rings 80 --size 1 1 100 --order 12 --groups 16 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02m_hex_s12_16g.json 2> gpurun_out/r02m_hex_s12_16g.err; show r02m_hex_s12_16g
