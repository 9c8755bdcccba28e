#!/bin/bash
# round-2 session 3, final tree, N GPUs: multi-GPU tests + the bench line at N with the tp_cfg4 / dp_cfg5 legs
N=${1:-2}
mkdir -p gpurun_out
[ "$N" = "8" ] && nvidia-smi topo -m > gpurun_out/s3f_topo_n$N.txt 2>&1
python -m pytest tests/test_gpu_gemm_allreduce.py tests/test_gpu_pdl.py -q -m gpu > gpurun_out/s3f_tests_n$N.log 2>&1
tail -2 gpurun_out/s3f_tests_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s3f_bench_n$N.json 2> gpurun_out/s3f_bench_n$N.err
echo "bench rc=$?"; tail -c 500 gpurun_out/s3f_bench_n$N.err; head -c 300 gpurun_out/s3f_bench_n$N.json
