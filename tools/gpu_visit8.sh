#!/bin/bash
# bench.py across the BASELINE.json configs on one GPU + reference / restatement arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "pixel or l1_loss_map" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/bench_c1.log 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"
timeout 900 python bench.py --config 5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --impl restatement --steps 3 --warmup 1 > gpurun_out/bench_restate.log 2> gpurun_out/bench_restate.err; echo "restate rc=$?"
for f in c3 c2 c1 c5 ref restate; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.log").read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("value","ms_per_step","e2e","roofline","stage_ms","value_fwd_bwd","dense_target","cuda_baseline","cpu_baseline","stats")}
    if keep.get("roofline"): keep["roofline"]={k:keep["roofline"][k] for k in ("kernel","frac","avg_launch_ms")}
    if keep.get("cuda_baseline"): keep["cuda_baseline"]={k:v for k,v in keep["cuda_baseline"].items() if k!="what"}
    print("$f", json.dumps(keep))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/bench_$f.err").read()[-2500:])
PY
done
