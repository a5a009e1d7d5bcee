#!/bin/bash
# usage: tools/quick_bench.sh tag  -> gpu tests (parity) + configs[2] and configs[1] device-resident numbers
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for c in 2 1; do
python bench.py --config $c --steps 10 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/${tag}_cfg$c.json 2> gpurun_out/${tag}_cfg$c.err || tail -5 gpurun_out/${tag}_cfg$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_cfg$c.json"))
print("cfg$c value %.0f ms/step %.3f e2e %.0f states %.0f conv %.4f kkt %.2e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("e2e_states",{}).get("value",0), d["converged_frac"], d["kkt_max"]))
PY
done
