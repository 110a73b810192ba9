#!/bin/bash
# Soak: many steps over all 64 cameras on every config (rare barrier-protocol hangs would trap here).
mkdir -p gpurun_out
for spec in "3 1500" "5 120" "2 1500" "1 1500"; do
  set -- $spec
  timeout 600 python bench.py --config $1 --steps $2 --warmup 3 --lean > gpurun_out/soak_c$1.log 2> gpurun_out/soak_c$1.err
  rc=$?
  python - $1 $rc <<'P'
import json,sys
c,rc=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f"gpurun_out/soak_c{c}.log").read().strip().splitlines()[-1])
    print("config",c,"rc",rc,"steps",d["steps"],round(d["value"],1),"views/s",round(d["ms_per_step"],3),"ms/step max_step",d["stats"].get("max_step_ms"))
except Exception as e:
    print("config",c,"rc",rc,"FAILED",e); print(open(f"gpurun_out/soak_c{c}.err").read()[-600:])
P
done
