#!/bin/bash
# Round-2 visit 1: full GPU parity suite, forward variant A/B bench, v3 role timeline.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
for V in 2 3; do
  GAGS_B200_FWD_VARIANT=$V timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/bench_v$V.log 2> gpurun_out/bench_v$V.err
  echo "bench v$V rc=$?" >> gpurun_out/bench_v$V.err
done
timeout 300 python tools/tc_timeline.py > gpurun_out/timeline_v3.txt 2>&1
tail -8 gpurun_out/pytest_gpu.log
for V in 2 3; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_v$V.log").read().strip().splitlines()[-1])
    print("v$V", round(d["value"],1), "views/s", {k: round(v,3) for k,v in d["stage_ms"].items()}, d["roofline"]["frac"])
except Exception as e:
    print("v$V failed", e); print(open("gpurun_out/bench_v$V.err").read()[-1500:])
PY
done
