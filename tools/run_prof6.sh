#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
{
for s in d4 d2 d1; do timeout 300 python tools/gpu_prof6.py pile64 4096 $s; done
timeout 300 python tools/gpu_prof6.py stack16 4096 d4
timeout 300 python tools/gpu_prof6.py stack16 4096 d2
} > gpurun_out/prof6_b.log 2>&1
cat gpurun_out/prof6_b.log
