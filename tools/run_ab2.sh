#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for l in vA vC nopause; do
  export ODEB_LIB_DIR=/root/repo/ode_b200/variants/$l
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ab_$l.csv python tools/profile_scene.py stack 4096 160 6 > /dev/null 2>&1
  echo "== $l"; python tools/launch_summary.py gpurun_out/ab_$l.csv 2
done
