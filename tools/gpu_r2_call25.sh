#!/bin/bash
# round 2: L2 eviction hints on the psi-row stores of the flow kernel (lanes streamed, edge copies kept): A/B on C4 and
# on the hexagonal lattice, DRAM bytes of the kernel with the hints, parity subset
mkdir -p gpurun_out
show() {
python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    p = d["step_phases_ms"]
    print(sys.argv[1], "ms/step %.3f kernel %.3f (frac %.3f) layout passes %.3f keff %s" % (d["ms_per_step"], p["sweep kernel alone"], d["roofline"]["frac"], p["shear + sweep + un-shear"] - p["sweep kernel alone"], d["config"]["keff_after_steps"]))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open("gpurun_out/%s.err" % sys.argv[1]).read()[-2000:])
PY
}
for rep in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02h_c4_hint$rep.json 2> gpurun_out/r02h_c4_hint$rep.err; show r02h_c4_hint$rep
PAMPA_SN_DBG=65536 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02h_c4_plain$rep.json 2> gpurun_out/r02h_c4_plain$rep.err; show r02h_c4_plain$rep
done
timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02h_hex_hint.json 2> gpurun_out/r02h_hex_hint.err; show r02h_hex_hint
PAMPA_SN_DBG=65536 timeout 300 python bench.py --mesh hex --no-cpu-baseline --no-e2e --no-solve > gpurun_out/r02h_hex_plain.json 2> gpurun_out/r02h_hex_plain.err; show r02h_hex_plain
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sn_sweep_flow -c 3 --csv \
   --log-file gpurun_out/r02h_flow_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches.log 2>&1
grep -v "^==" gpurun_out/r02h_flow_dram.csv | cut -d, -f5,13- | tail -12
( timeout 600 python -u -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "schedule or single_sweep or fused" ) > gpurun_out/r02_pytest_sel.log 2>&1
tail -3 gpurun_out/r02_pytest_sel.log; grep -B5 -A30 "^E " gpurun_out/r02_pytest_sel.log | head -60
