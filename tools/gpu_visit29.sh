#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sign_operand" --tb=short 2>&1 | tail -4 | cut -c1-300
GAGS_B200_BWD_SIGN=2 timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "multi_job or benchmark_config" --tb=short 2>&1 | tail -3 | cut -c1-300
for A in 2 1; do
GAGS_B200_BWD_SIGN=$A timeout 600 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/bench_sgn$A.log 2> gpurun_out/bench_sgn$A.err
python - $A <<'P'
import json,sys
a=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_sgn{a}.log").read().strip().splitlines()[-1])
    print("sign",a, round(d["value"],1),"views/s", round(d["ms_per_step"],3),"ms/step", {k:round(v,3) for k,v in d["stage_ms"].items() if v>0.05}, "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/bench_sgn{a}.err").read()[-800:])
P
done
