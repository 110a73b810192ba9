#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for A in 1 0; do
GAGS_B200_ASYNC_ADAM=$A timeout 600 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/bench_async$A.log 2> gpurun_out/bench_async$A.err
python - $A <<'P'
import json,sys
a=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_async{a}.log").read().strip().splitlines()[-1])
    print("async",a, round(d["value"],1),"views/s", round(d["ms_per_step"],3),"ms/step", {k:round(v,3) for k,v in d["stage_ms"].items() if v>0.05})
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/bench_async{a}.err").read()[-800:])
P
done
