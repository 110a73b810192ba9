#!/bin/bash
# Round-2f single-GPU bench lines (final kernels): config 3 full line, reference arm, configs 1 / 2 / 5.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --config 5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/bench_c1.log 2> gpurun_out/bench_c1.err
cat gpurun_out/bench.log | cut -c1-6000; tail -3 gpurun_out/bench.err
for c in c1 c2 c5; do python - $c <<'P'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{c}.log").read().strip().splitlines()[-1])
    print(c, round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms/step e2e", d.get("e2e") and round(d["e2e"]["value"],1), {k:round(v,3) for k,v in (d.get("stage_ms") or {}).items() if v>0.05}, "roofline", d.get("roofline",{}) and round(d["roofline"]["frac"],3), d.get("cuda_baseline",{}) and d["cuda_baseline"].get("value"))
except Exception as e: print(c,"failed",e)
P
done
