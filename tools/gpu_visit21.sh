#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python tools/two_pass_probe.py 3 > gpurun_out/two_pass.txt 2>&1
tail -5 gpurun_out/two_pass.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_lazy.log 2> gpurun_out/bench_lazy.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_lazy.log").read().strip().splitlines()[-1])
print(round(d["value"],1),"views/s", round(d["ms_per_step"],3),"ms/step; e2e", d["e2e"] and round(d["e2e"]["value"],1))
print({k:round(v,3) for k,v in d["stage_ms"].items()})
P
tail -3 gpurun_out/bench_lazy.err
