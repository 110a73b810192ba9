#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lazy_training" --tb=short 2>&1 | grep -v "^E    +\|^E  +\|^E     +" | tail -30 | cut -c1-250
