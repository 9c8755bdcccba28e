#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_config_scale.py tests/test_gpu_attention_graph.py tests/test_gpu_pdl.py tests/test_gpu_attention_sm100.py tests/test_gpu_regressions.py tests/test_gpu_runtime.py -x -q -m gpu 2>&1 | tail -4
for f in 1 0; do echo "== FOLD=$f"; MOJO_B200_DECODE_FOLD=$f python tools/bench_decode_small.py 2>&1 | tail -4; done
MOJO_B200_LIB=mojo_opset_b200/libmojo_b200_dtr.so python tools/decode_trace.py 1 32768
python bench.py --steps 20 --warmup 5 --sustain-s 0 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'])"
