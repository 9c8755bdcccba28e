#!/bin/bash
# TP step A/B at N GPUs: PDL / pair / prefetch variants of the fused GEMM + all-reduce
N=${1:-2}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 \
    tools/bench_tp_step.py --steps 20 --warmup 5 2> gpurun_out/s3_tp_n${N}_$name.err | tail -1 | tee -a gpurun_out/s3_tp_n$N.jsonl | cut -c1-420
}
rm -f gpurun_out/s3_tp_n$N.jsonl
run base X=1
run nopdl MOJO_B200_PDL=0
run nopair MOJO_B200_GAR_PAIR=0
run pf12 MOJO_B200_GAR_PREFETCH=12
run pfall MOJO_B200_GAR_PREFETCH=100000
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 tools/bench_gemm_allreduce.py --check --steps 20 --warmup 5 2>gpurun_out/s3_gar_n$N.err | tail -3 | tee gpurun_out/s3_gar_n$N.txt | cut -c1-900
