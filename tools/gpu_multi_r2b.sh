#!/bin/bash
# Multi-GPU visit (round 2b): SparsePeerAdam / PeerAdam checks + bench.py under torchrun at $NG ranks,
# row-sparse exchange (default) and the dense fused exchange.
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for NVLS in 0 1; do
  GAGS_B200_NVLS=$NVLS timeout 300 $TR --master-port 29515 tests/multi_gpu/sparse_peer_adam_check.py > gpurun_out/sparse_check_${NG}gpu_nvls$NVLS.log 2>&1
  echo "sparse check (NVLS=$NVLS) rc=$?"; grep -h "rank\|Error\|error" gpurun_out/sparse_check_${NG}gpu_nvls$NVLS.log | tail -12
done
GAGS_B200_NVLS=auto timeout 300 $TR --master-port 29515 tests/multi_gpu/peer_adam_check.py > gpurun_out/peer_check_${NG}gpu.log 2>&1
echo "dense peer check rc=$?"; grep -h "rank\|Error\|error" gpurun_out/peer_check_${NG}gpu.log | tail -8
export GAGS_B200_PEER_TIMING=1
for MODE in ${MODES:-sparse dense}; do
  FLAG=""; [ "$MODE" = "dense" ] && FLAG="--dense-exchange"
  [ "$MODE" = "nccl" ] && FLAG="--nccl-allreduce"
  timeout 600 $TR --master-port 29517 bench.py --gpus $NG --steps ${STEPS:-20} --warmup 3 $FLAG \
    --no-cpu-baseline > gpurun_out/bench_${NG}gpu_$MODE.log 2> gpurun_out/bench_${NG}gpu_$MODE.err
  echo "$MODE rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_$MODE.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'],'gpus', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['value'],1) if d.get('e2e') else None, d['config'].get('grad_exchange'), d.get('stats',{}).get('peer_step_ms'), d.get('exchange_check'))"
  grep -h "PeerAdam\|Error" gpurun_out/bench_${NG}gpu_$MODE.err | tail -3
done
