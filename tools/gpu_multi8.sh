#!/bin/bash
# One 8-GPU visit (expensive): bench.py with the default exchange (NVLS multimem at 8 ranks), then
# with unicast peer loads/stores.
NG=${NG:-8}
mkdir -p gpurun_out
export GAGS_B200_PEER_TIMING=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for MODE in ${MODES:-auto unicast}; do
  [ "$MODE" = "unicast" ] && export GAGS_B200_NVLS=0
  [ "$MODE" = "nvls" ] && export GAGS_B200_NVLS=1
  timeout 400 $TR --master-port 29517 bench.py --gpus $NG --steps ${STEPS:-10} --warmup 3 \
    --no-cpu-baseline > gpurun_out/bench_${NG}gpu_$MODE.log 2> gpurun_out/bench_${NG}gpu_$MODE.err
  echo "$MODE rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_$MODE.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'],'gpus', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['value'],1) if d.get('e2e') else None, d['config'].get('grad_exchange'), d.get('stats',{}).get('peer_step_ms'))"
  grep -h "PeerAdam\|Error" gpurun_out/bench_${NG}gpu_$MODE.err | tail -3
done
