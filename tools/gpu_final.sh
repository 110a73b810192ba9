#!/bin/bash
# Final-tree check: GPU suite, smoke, the default bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2 | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print(round(d["value"],1),"views/s", round(d["ms_per_step"],3),"ms/step e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], "clocks", d["clocks"])
P
