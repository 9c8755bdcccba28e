#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_dit_ops.py tests/test_gpu_norm_rope_store.py tests/test_gpu_elementwise_large.py tests/test_gpu_dit_block.py tests/test_gpu_hf_patch.py tests/test_gpu_runtime.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/bench_ops.py --skip-decode --json gpurun_out/s3_ops_roofline.json 2>&1 | grep "8192\|262144"
echo "== old norm_rope_store"
MOJO_B200_LIB=mojo_opset_b200/libmojo_b200_nrsold.so timeout 600 python tools/bench_ops.py --skip-decode 2>&1 | grep "fused"
