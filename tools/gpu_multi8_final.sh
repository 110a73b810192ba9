#!/bin/bash
# Final multi-GPU visit of round 2: config 3 (default grid and a small exchange grid) and config 5 at
# $NG ranks with the row-sparse exchange + lazily evaluated Adam.
NG=${NG:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/gpus_${NG}.txt 2>&1
run() { # tag cfg steps env...
  tag=$1; cfg=$2; st=$3; shift 3
  env "$@" timeout 600 $TR --master-port 29517 bench.py --gpus $NG --config $cfg --steps $st --warmup 3 \
    --no-cpu-baseline > gpurun_out/bench_${NG}gpu_$tag.log 2> gpurun_out/bench_${NG}gpu_$tag.err
  echo "$tag rc=$?"; tail -1 gpurun_out/bench_${NG}gpu_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print(d['n_gpus'],'gpus', round(d['value'],1),'views/s', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['value'],1) if d.get('e2e') else None)
    print(' stage_ms', {k: round(v,3) for k,v in d['stage_ms'].items() if v > 0.02})
    print(' exchange_check', {k: d['exchange_check'][k] for k in ('rel_err','replicas_identical','grad_sum_vs_single_process_rel_err','loss_sum_matches_single_process')} if d.get('exchange_check') else None)
    print(' peer', d.get('stats',{}).get('peer_step_ms'), d.get('stats',{}).get('peer_by_rank',{}).get('wait_slowest_rank_ms'))
except Exception as e: print('parse failed', e)
"
  grep -h "Error\|error" gpurun_out/bench_${NG}gpu_$tag.err | tail -3
}
run c3 3 20 A=1
run c3_grid2 3 20 GAGS_B200_PEER_GRID=2
run c5 5 8 A=1
