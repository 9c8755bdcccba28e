#!/bin/bash
for i in 1 2; do
for lib in mojo_opset_b200/libmojo_b200.so mojo_opset_b200/libmojo_b200_dold.so; do
MOJO_B200_LIB=$lib python bench.py --steps 20 --warmup 5 --sustain-s 0 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'value', round(d['value']), d['ms_per_step'], 'decode us', d['roofline']['us_per_launch'])"
done; done
python tools/bench_decode_small.py 2>&1 | tail -4
python -m pytest tests/test_gpu_golden.py tests/test_gpu_config_scale.py tests/test_gpu_attention_graph.py tests/test_gpu_pdl.py tests/test_gpu_attention_sm100.py tests/test_gpu_regressions.py tests/test_gpu_runtime.py -x -q -m gpu 2>&1 | tail -2
