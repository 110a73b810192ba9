#!/bin/bash
# Round-2e single-GPU evidence visit: parity tests, smoke, the bench lines of every arm and config,
# ncu launch list + full capture of the blend kernels of one steady-state step.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/bench_c1.log 2> gpurun_out/bench_c1.err
timeout 600 python bench.py --config 5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --lean \
    > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 27 -c 3 \
    -f -o gpurun_out/prof_blend3 python bench.py --steps 2 --warmup 1 --lean \
    > gpurun_out/ncu_full.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-4500; tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_ref.log
for c in c1 c2 c5; do python - $c <<'P'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{c}.log").read().strip().splitlines()[-1])
    print(c, round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms/step", {k:round(v,3) for k,v in (d.get("stage_ms") or {}).items() if v>0.05}, "roofline", d.get("roofline",{}) and round(d["roofline"]["frac"],3))
except Exception as e: print(c,"failed",e)
P
done
