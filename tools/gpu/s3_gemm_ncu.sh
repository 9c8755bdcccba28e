#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_allreduce|gemm|Kernel|cutlass|nvjet' -c 12 -o gpurun_out/s3_gemm_prof -f python tools/profile_gemm.py > gpurun_out/s3_gemm_ncu.log 2>&1
echo rc=$?; tail -5 gpurun_out/s3_gemm_ncu.log; ls -la gpurun_out/s3_gemm_prof.ncu-rep
