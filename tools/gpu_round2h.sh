#!/bin/bash
# Final visit of round 2: the whole GPU suite, smoke, the default bench line, one A/B of the backward
# grid (1 vs 2 persistent CTAs per SM).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
GAGS_B200_BWD_CTAS=1 timeout 600 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/bench_bwd1.log 2> gpurun_out/bench_bwd1.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log | cut -c1-200; tail -2 gpurun_out/bench.err
for n in bench bench_bwd1; do python - $n <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1])
    print(n, round(d["value"],1),"views/s", round(d["ms_per_step"],3),"ms/step e2e", d.get("e2e") and round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items() if v>0.05}, "frac", round(d["roofline"]["frac"],3), {k:v for k,v in d["stats"].items() if k.startswith(("gauss","k_eff","n_"))})
except Exception as e: print(n,"failed",e)
P
done
