#!/bin/bash
# round-2 GPU session 2 (N GPUs): multi-GPU tests + bench line at N with the TP cfg4 / cfg5 legs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s2_topo_n$N.txt 2>&1
python -m pytest tests/test_gpu_gemm_allreduce.py tests/test_gpu_regressions.py -q -m gpu > gpurun_out/s2_tests_n$N.log 2>&1
tail -3 gpurun_out/s2_tests_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s2_bench_n$N.json 2> gpurun_out/s2_bench_n$N.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/s2_bench_n$N.err; head -c 300 gpurun_out/s2_bench_n$N.json
