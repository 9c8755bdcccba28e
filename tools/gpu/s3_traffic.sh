#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_regressions.py tests/test_gpu_golden.py tests/test_gpu_attention_graph.py -x -q -m gpu 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:paged_decode_mma -s 4 -c 2 -f -o gpurun_out/r2s3_decode \
  python bench.py --steps 6 --warmup 4 --repeats 1 --sustain-s 0 --no-extra --no-cpu-baseline > gpurun_out/s3f_ncu_decode.log 2>&1
echo ncu rc=$?
