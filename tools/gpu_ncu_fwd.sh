#!/bin/bash
# ncu --set full capture of the forward blend kernel only (one launch), variant from $V (default 3)
mkdir -p gpurun_out
export GAGS_B200_FWD_VARIANT=${V:-3}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_fwd_tc -s 8 -c 1 \
  -f -o gpurun_out/prof_fwd_v${V:-3} python bench.py --steps 2 --warmup 1 --lean \
  > gpurun_out/ncu_fwd_v${V:-3}.log 2>&1
tail -3 gpurun_out/ncu_fwd_v${V:-3}.log
