#!/bin/bash
# Multi-GPU visit: PeerAdam check + bench.py under torchrun at $NG ranks (fused peer step and, with
# NCCL=1, the NCCL all-reduce baseline).
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29515 tests/multi_gpu/peer_adam_check.py > gpurun_out/peer_check_${NG}gpu.log 2>&1
echo "peer check rc=$?"; grep -h "rank\|Error\|error" gpurun_out/peer_check_${NG}gpu.log | tail -12
for MODE in ${MODES:-peer nccl}; do
  FLAG=""; [ "$MODE" = "nccl" ] && FLAG="--nccl-allreduce"
  timeout 600 $TR --master-port 29517 bench.py --gpus $NG --steps ${STEPS:-10} --warmup 3 $FLAG \
    --no-cpu-baseline > gpurun_out/bench_${NG}gpu_$MODE.log 2> gpurun_out/bench_${NG}gpu_$MODE.err
  echo "$MODE rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_$MODE.log | cut -c1-330
  grep -h "PeerAdam\|Error" gpurun_out/bench_${NG}gpu_$MODE.err | tail -3
done
