#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
{
for s in d4; do timeout 400 python tools/gpu_solve6_check.py parity $s single || echo "PARITY FAILED/timeout $s single rc=$?"; done
for s in d8 d4 p4; do timeout 300 python tools/gpu_solve6_check.py time pile64 4096 $s || echo "TIME FAILED $s"; done
for s in d4; do timeout 300 python tools/gpu_prof6.py pile64 4096 $s; done
for s in d4 d2 d1 p4; do timeout 400 python tools/gpu_solve6_check.py time pile64s 4096 $s || echo "TIME FAILED $s"; done
for s in d4 d1; do timeout 400 python tools/gpu_prof6.py pile64s 4096 $s; done
for s in d4; do timeout 300 python tools/gpu_solve6_check.py time stack16 4096 $s 50 || echo "TIME FAILED $s"; done
} > gpurun_out/solve6_f.log 2>&1
grep -v "bit-exact" gpurun_out/solve6_f.log | tail -40; grep -c "bit-exact" gpurun_out/solve6_f.log
