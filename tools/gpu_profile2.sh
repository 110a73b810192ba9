#!/bin/bash
# ncu captures: launch list + full capture of the three blend kernels (weights pass, blend pass,
# cached backward) of one steady-state step.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --lean \
  > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 27 -c 3 \
  -f -o gpurun_out/prof_blend3 python bench.py --steps 2 --warmup 1 --lean \
  > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/prof_blend3.ncu-rep
