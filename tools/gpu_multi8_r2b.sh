#!/bin/bash
# Round-2b multi-GPU visit: SparsePeerAdam check at $NG ranks (NVLS form), config 3 (and 5) bench lines
# with the row-sparse exchange + lazily evaluated Adam, the multi-GPU pytest.  Lands in gpurun_out/.
NG=${NG:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/gpus_${NG}.txt 2>&1
timeout 300 $TR --master-port 29515 tests/multi_gpu/sparse_peer_adam_check.py > gpurun_out/sparse_check_${NG}gpu.log 2>&1
echo "sparse check rc=$?"; grep -h "rank 0\|Error\|error" gpurun_out/sparse_check_${NG}gpu.log | tail -6 | cut -c1-250
for CFG in ${CONFIGS:-3 5}; do
  ST=20; [ "$CFG" = "5" ] && ST=8
  timeout 900 $TR --master-port 29517 bench.py --gpus $NG --config $CFG --steps $ST --warmup 3 \
    --no-cpu-baseline $EXTRA > gpurun_out/bench_${NG}gpu_c$CFG.log 2> gpurun_out/bench_${NG}gpu_c$CFG.err
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
