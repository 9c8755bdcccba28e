#!/bin/bash
# compute-sanitizer memcheck over the kernels changed in round-2 session 3 (PDL chain, split fold, packed norms, store, RoPE, line protocol at world 1 / 2)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_regressions.py tests/test_gpu_golden.py tests/test_gpu_norm_rope_store.py tests/test_gpu_pdl.py tests/test_gpu_dit_ops.py -x -q -m gpu > gpurun_out/s3_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/s3_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm_allreduce.py -x -q -m gpu -k "single_rank or (virtual and 2-)" > gpurun_out/s3_memcheck_gar.log 2>&1
echo "memcheck gar rc=$?"; tail -6 gpurun_out/s3_memcheck_gar.log
