#!/bin/bash
# Quick GPU-box visit: parity tests + short bench (+ optional forward timeline), no ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
if [ "${TIMELINE:-0}" = "1" ]; then timeout 300 python tools/tc_timeline.py > gpurun_out/timeline.txt 2>&1; fi
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
