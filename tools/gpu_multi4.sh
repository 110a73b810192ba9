#!/bin/bash
NG=${NG:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${NG}gpu_c3.log 2> gpurun_out/bench_${NG}gpu_c3.err
echo "rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_c3.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print(d['n_gpus'],'gpus', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['value'],1) if d.get('e2e') else None, d['config']['grad_exchange'][:60])
    print(' exchange_check', {k: d['exchange_check'][k] for k in ('rel_err','replicas_identical','grad_sum_vs_single_process_rel_err','loss_sum_matches_single_process')})
    print(' peer', d['stats'].get('peer_step_ms'))
except Exception as e: print('parse failed', e)
"
grep -h "Error\|error" gpurun_out/bench_${NG}gpu_c3.err | tail -3
