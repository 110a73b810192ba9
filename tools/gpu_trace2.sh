#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/trace_step_dp.py > gpurun_out/trace_dp2_sparse.txt 2>&1
tail -75 gpurun_out/trace_dp2_sparse.txt | cut -c1-110
