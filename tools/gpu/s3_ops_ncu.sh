#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'norm_rope_store|apply_rope_slice|store_table|act_kernel' -s 4 -c 5 -f -o gpurun_out/r2s3_ops python tools/profile_ops.py > gpurun_out/s3_ops_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/s3_ops_ncu.log
