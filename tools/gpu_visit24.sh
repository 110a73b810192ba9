#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1])
    print(n, round(d["value"],1),"views/s", round(d["ms_per_step"],3),"ms/step", {k:round(v,3) for k,v in d["stage_ms"].items() if v>0.05})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-600:])
P
}
run default A=1
run lookahead GAGS_B200_LOOKAHEAD=1
run lazygrid16 GAGS_B200_LAZY_GRID=16
run lazygrid4 GAGS_B200_LAZY_GRID=4
