#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/two_pass_probe.py 3 > gpurun_out/two_pass.txt 2>&1
tail -8 gpurun_out/two_pass.txt
