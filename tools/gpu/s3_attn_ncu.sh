#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_sm100 -c 4 -f -o gpurun_out/r2s3_attn python tools/profile_attn.py > gpurun_out/s3_attn_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/s3_attn_ncu.log
