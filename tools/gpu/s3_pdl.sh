#!/bin/bash
# round-2 session 3: PDL on/off A/B (headline step, small-batch decode), parity subset
mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden.py tests/test_gpu_config_scale.py tests/test_gpu_gemm_allreduce.py tests/test_gpu_norm_rope_store.py tests/test_gpu_attention_graph.py tests/test_gpu_runtime.py -x -q -m gpu > gpurun_out/s3_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/s3_tests.log
for pdl in 1 0; do
  echo "== PDL=$pdl"
  MOJO_B200_PDL=$pdl python tools/bench_decode_small.py 2>&1 | tail -5
  MOJO_B200_PDL=$pdl python bench.py --steps 20 --warmup 5 --sustain-s 0 --no-extra --no-cpu-baseline > gpurun_out/s3_bench_pdl$pdl.json 2> gpurun_out/s3_bench_pdl$pdl.err
  echo "bench rc=$?"; tail -c 600 gpurun_out/s3_bench_pdl$pdl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s3_bench_pdl$pdl.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "eager", d.get("eager"), "roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
PY
done
