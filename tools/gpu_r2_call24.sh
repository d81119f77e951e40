#!/bin/bash
# round 2: hexagonal lattice with chunks of up to 6 directions (classes of 6 directions in one chunk), and the
# reference (CPU) arm of the bench on the GPU box's host cores
mkdir -p gpurun_out
show() {
python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    p = d["step_phases_ms"]
    print(sys.argv[1], "ms/step %.3f kernel %.3f (frac %.3f) layout passes %.3f launches/step %s keff %s" % (d["ms_per_step"], p["sweep kernel alone"], d["roofline"]["frac"], p["shear + sweep + un-shear"] - p["sweep kernel alone"], d["roofline"]["launches_per_step"], d["config"]["keff_after_steps"]))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/%s.err" % sys.argv[1]).read()[-2000:])
PY
}
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02g_hex_dt4.json 2> gpurun_out/r02g_hex_dt4.err; show r02g_hex_dt4
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve --opts '{"dt_max": 6}' > gpurun_out/r02g_hex_dt6.json 2> gpurun_out/r02g_hex_dt6.err; show r02g_hex_dt6
timeout 300 python bench.py --mesh hex --rings 80 --size 1 1 100 --order 12 --groups 16 --no-cpu-baseline --no-e2e --no-solve --opts '{"dt_max": 5}' > gpurun_out/r02g_hex_s12_dt5.json 2> gpurun_out/r02g_hex_s12_dt5.err; show r02g_hex_s12_dt5
timeout 300 python bench.py --mesh hex --rings 80 --size 1 1 100 --order 12 --groups 16 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02g_hex_s12_dt4.json 2> gpurun_out/r02g_hex_s12_dt4.err; show r02g_hex_s12_dt4
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02g_reference.json 2> gpurun_out/r02g_reference.err
tail -c 1500 gpurun_out/r02g_reference.json; tail -3 gpurun_out/r02g_reference.err
