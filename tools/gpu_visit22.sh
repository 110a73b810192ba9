#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_sparse_autograd.py 2>&1 | tail -20 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
