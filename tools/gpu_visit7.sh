#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
for V in 3 2; do
  GAGS_B200_FWD_VARIANT=$V timeout 600 python bench.py --steps 10 --warmup 3 --lean \
    > gpurun_out/bench_v$V.log 2> gpurun_out/bench_v$V.err
  echo "bench v$V rc=$?" >> gpurun_out/bench_v$V.err
done
V=3 bash tools/gpu_ncu_fwd.sh
timeout 300 python tools/tc_timeline.py > gpurun_out/timeline_v3.txt 2>&1
tail -12 gpurun_out/pytest_gpu.log
for V in 3 2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_v$V.log").read().strip().splitlines()[-1])
    print("v$V", round(d["value"],1), "views/s", {k: round(v,3) for k,v in d["stage_ms"].items()}, d["roofline"]["frac"])
except Exception as e:
    print("v$V failed", e); print(open("gpurun_out/bench_v$V.err").read()[-1500:])
PY
done
