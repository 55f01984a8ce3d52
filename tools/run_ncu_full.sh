#!/bin/bash
# ncu --set full captures of the shipped solver kernels (one launch each), summaries under gpurun_out/
cd /root/repo; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve6 -c 1 -o gpurun_out/r2_solve6_pile64 -f python tools/profile_scene.py pile64 4096 156 1 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_lwt_phase -s 1 -c 1 -o gpurun_out/r2_lwt_phase -f python tools/profile_scene.py wall 1 8 1 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve -c 2 -o gpurun_out/r2_hybrid_stack -f python tools/profile_scene.py stack 4096 170 1 > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_islands -c 1 -o gpurun_out/r2_islands_pile64 -f python tools/profile_scene.py pile64 4096 156 1 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
