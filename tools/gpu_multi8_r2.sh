#!/bin/bash
# Round-2 8-GPU visit: exchange-kernel rate sweep, config 3 and config 5 bench lines at 8 ranks,
# the two-rank PeerAdam pytest.  Everything lands in gpurun_out/.
NG=${NG:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/gpus_${NG}.txt 2>&1
nproc >> gpurun_out/gpus_${NG}.txt; free -g | head -2 >> gpurun_out/gpus_${NG}.txt
if [ "${SWEEP:-1}" = "1" ]; then
  timeout 400 $TR --master-port 29511 tools/peer_rate.py > gpurun_out/peer_rate_${NG}gpu.log 2>&1
  echo "peer_rate rc=$?"; grep -h "ranks" gpurun_out/peer_rate_${NG}gpu.log | sort -t= -k2 | tail -40
fi
[ -n "$PEER_GRID" ] && export GAGS_B200_PEER_GRID=$PEER_GRID
[ -n "$PEER_UNROLL" ] && export GAGS_B200_PEER_UNROLL=$PEER_UNROLL
for CFG in ${CONFIGS:-3 5}; do
  ST=20; [ "$CFG" = "5" ] && ST=8
  timeout 900 $TR --master-port 29517 bench.py --gpus $NG --config $CFG --steps $ST --warmup 3 \
    --no-cpu-baseline > gpurun_out/bench_${NG}gpu_c$CFG.log 2> gpurun_out/bench_${NG}gpu_c$CFG.err
  echo "config $CFG rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_c$CFG.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print(d['n_gpus'],'gpus', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['value'],1) if d.get('e2e') else None, d['config'].get('grad_exchange'))
    print(' stage_ms', {k: round(v,3) for k,v in d['stage_ms'].items()})
    print(' exchange_check', d.get('exchange_check'))
    print(' stats', d.get('stats'))
except Exception as e: print('parse failed', e)
"
  grep -h "PeerAdam\|Error\|error" gpurun_out/bench_${NG}gpu_c$CFG.err | tail -5
done
if [ "${PYTEST:-1}" = "1" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
fi
