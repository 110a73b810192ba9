#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_pass or lazy_training" --tb=short 2>&1 | grep -v "^E    +\|^E  +\|^E     +" | tail -15 | cut -c1-250
timeout 200 python tools/two_pass_probe.py 3 > gpurun_out/two_pass.txt 2>&1
tail -5 gpurun_out/two_pass.txt
