#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
export ODEB_LW_SWEEP=4
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_lwc_sweep -s 20 -c 1 -o gpurun_out/r2_lwc_sweep4 -f python tools/profile_scene.py wall 1 8 1 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/r2_lwc_sweep4.ncu-rep
