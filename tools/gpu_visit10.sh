#!/bin/bash
# 2-GPU visit: DP step trace + bench at 2 ranks (k = 1 and k = 2)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tools/trace_step_dp.py > gpurun_out/trace_dp2.txt 2>&1; echo "trace rc=$?"
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_c3.log 2> gpurun_out/bench_2gpu_c3.err; echo "bench2 rc=$?"
tail -1 gpurun_out/bench_2gpu_c3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'],'gpus', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step', {k: round(v,3) for k,v in d['stage_ms'].items()}, d.get('exchange_check',{}).get('rel_err'))"
grep -n "adam_peer\|adam_multicast" gpurun_out/trace_dp2.txt | head -3
