#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for sc in "pile64 4096 155" "ragdoll 2048 60" "stack 4096 160" "chain 65536 40"; do
  set -- $sc
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_$1.csv python tools/profile_scene.py $1 $2 $3 3 > gpurun_out/prof_$1.log 2>&1
  echo "== $1 $2 worlds, 3 steps"; python tools/launch_summary.py gpurun_out/r2_launches_$1.csv 16 2>&1 | tee gpurun_out/r2_launches_$1.txt
done
