#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_tc.log 2> gpurun_out/bench_c2_tc.err; echo "c2 tc rc=$?"
GAGS_B200_BLEND_IMPL=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c2_simt.log 2> gpurun_out/bench_c2_simt.err; echo "c2 simt rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blend_forward or feature_backward_matches or wide_blend or full_backward" 2>&1 | tail -3
for f in c2_tc c2_simt; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.log").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "views/s e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["stage_ms"].items()}, (d.get("cuda_baseline") or {}).get("value"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/bench_$f.err").read()[-1500:])
PY
done
