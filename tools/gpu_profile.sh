#!/bin/bash
# ncu captures only: launch list + full capture of the blend kernels (one bench process each).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline \
  > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 2 -c 2 \
  -f -o gpurun_out/prof_blend python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline \
  > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
