#!/bin/bash
# round-2 session 3, final tree, one B200: whole GPU suite, bench line, reference arm, decode ncu capture, launch list, smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s3f_smi.txt 2>&1
python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/s3f_tests_all.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/s3f_tests_all.log
python bench.py > gpurun_out/s3f_bench_n1.json 2> gpurun_out/s3f_bench_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/s3f_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3f_bench_ref.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:paged_decode_mma -s 4 -c 2 -f -o gpurun_out/r2s3_decode \
  python bench.py --steps 6 --warmup 4 --repeats 1 --sustain-s 0 --no-extra --no-cpu-baseline > gpurun_out/s3f_ncu_decode.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2s3_launches_bench_cfg2.csv \
  python bench.py --steps 6 --warmup 4 --repeats 1 --sustain-s 0 --no-extra --no-cpu-baseline > gpurun_out/s3f_ncu_launches.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
