#!/bin/bash
# the bench scene (4096 x 16-box stacks) under every solver kernel family
cd /root/repo
for s in "" d4 d2 d8 p4 v4; do ODEB_SOLVER=$s python tools/gpu_ab_time.py stack 4096 256 | tail -1 | sed "s/^/solver=[$s] /"; done
