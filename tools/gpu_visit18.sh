#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tools/diag_full_bwd.py > gpurun_out/diag_full_bwd.txt 2>&1
cat gpurun_out/diag_full_bwd.txt | tail -24
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sparse.log 2> gpurun_out/bench_sparse.err
tail -c 3000 gpurun_out/bench_sparse.log; tail -5 gpurun_out/bench_sparse.err
GAGS_B200_SPARSE_ADAM=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/bench_dense.log 2> gpurun_out/bench_dense.err
tail -c 1500 gpurun_out/bench_dense.log
