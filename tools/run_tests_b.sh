#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
{
timeout 900 python tools/measure_tolerances.py
timeout 1700 python -m pytest tests -m gpu -x -q -k "golden_colliders or vs_compiled_reference or sampled_worlds or pile_1000 or walker or hybrid_mixed" 2>&1 | tail -30
} > gpurun_out/tests_b.log 2>&1
tail -60 gpurun_out/tests_b.log
