#!/bin/bash
# Round-2 single-GPU evidence visit: parity tests, smoke, the bench lines of every arm and config,
# ncu launch list + full capture of the two blend kernels.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --config 2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/bench_c1.log 2> gpurun_out/bench_c1.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --lean \
    > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 8 -c 2 \
    -f -o gpurun_out/prof_blend python bench.py --steps 2 --warmup 1 --lean \
    > gpurun_out/ncu_full.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-3000; tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_ref.log
