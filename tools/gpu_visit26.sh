#!/bin/bash
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg.log 2> gpurun_out/bench_dbg.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_dbg.log").read().strip().splitlines()[-1])
print(round(d["value"],1), d["dense_target"], d["value_fwd_bwd"])
P
