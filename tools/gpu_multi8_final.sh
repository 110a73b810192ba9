#!/bin/bash
# 8-GPU confirmation of the small exchange grids + views-per-step 2 lines
NG=${NG:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for K in 1 2; do
  timeout 600 $TR --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 --views-per-step $K \
    --no-cpu-baseline > gpurun_out/bench_${NG}gpu_k$K.log 2> gpurun_out/bench_${NG}gpu_k$K.err
  echo "k=$K rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_k$K.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print(d['n_gpus'],'gpus k=$K', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['value'],1) if d.get('e2e') else None)
    print(' stage_ms', {k: round(v,3) for k,v in d['stage_ms'].items()})
    print(' exchange_check', {k:v for k,v in (d.get('exchange_check') or {}).items() if k!='reference'})
except Exception as e: print('parse failed', e)
"
  grep -h "Error\|error" gpurun_out/bench_${NG}gpu_k$K.err | tail -3
done
timeout 300 python bench.py --steps 20 --warmup 3 --views-per-step 2 --lean > gpurun_out/bench_1gpu_k2.log 2>gpurun_out/bench_1gpu_k2.err
tail -1 gpurun_out/bench_1gpu_k2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1 gpu k=2', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step')"
timeout 300 python bench.py --steps 20 --warmup 3 --lean > gpurun_out/bench_1gpu_k1.log 2>gpurun_out/bench_1gpu_k1.err
tail -1 gpurun_out/bench_1gpu_k1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1 gpu k=1', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step')"
