#!/bin/bash
# TP step at N GPUs (tp_cfg4 leg only) + the stand-alone fused GEMM + all-reduce bench; extra args: variants to run
N=${1:-2}; shift
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 \
    tools/bench_tp_step.py --steps 20 --warmup 5 2> gpurun_out/s3_tp_n${N}_$name.err | tail -1 | tee -a gpurun_out/s3_tp_n$N.jsonl | cut -c1-420
}
rm -f gpurun_out/s3_tp_n$N.jsonl
run base X=1
for v in "$@"; do
  case $v in
    nopdl) run nopdl MOJO_B200_PDL=0;;
    nopair) run nopair MOJO_B200_GAR_PAIR=0;;
    two) run two MOJO_B200_GAR_MODE=two;;
    one) run one MOJO_B200_GAR_MODE=one;;
    nobench) exit 0;;
  esac
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 tools/bench_gemm_allreduce.py --check --steps 20 --warmup 5 2>gpurun_out/s3_gar_n$N.err | tail -3 | tee gpurun_out/s3_gar_n$N.txt | cut -c1-900
