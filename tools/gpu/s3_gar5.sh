#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_allreduce.py -x -q -m gpu 2>&1 | tail -3
{
echo "== pair k4096 warm"; MOJO_B200_LIB=mojo_opset_b200/libmojo_b200_tr.so python tools/gar_trace.py 256 8192 4096
echo "== pair k1024 cold"; MOJO_B200_LIB=mojo_opset_b200/libmojo_b200_tr.so python tools/gar_trace.py 256 8192 1024 --cold
} > gpurun_out/s3_gar_trace5.txt 2>&1
python tools/bench_gemm_own.py 2>&1 | tee gpurun_out/s3_gemm_own5.txt
