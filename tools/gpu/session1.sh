#!/bin/bash
# round-2 GPU session 1 (one B200): new parity tests, whole GPU suite, bench line, reference arm, decode ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
nproc >> gpurun_out/s1_smi.txt
python -m pytest tests/test_gpu_config_scale.py tests/test_gpu_regressions.py tests/test_gpu_reference_plugin.py -q -m gpu --durations=8 > gpurun_out/s1_tests_new.log 2>&1
tail -5 gpurun_out/s1_tests_new.log
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_config_scale.py --deselect tests/test_gpu_reference_plugin.py > gpurun_out/s1_tests_all.log 2>&1
tail -3 gpurun_out/s1_tests_all.log
python bench.py > gpurun_out/s1_bench_n1.json 2> gpurun_out/s1_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/s1_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s1_bench_ref.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:paged_decode_mma -s 4 -c 2 -o gpurun_out/r2_decode \
  python bench.py --steps 6 --warmup 4 --repeats 1 --sustain-s 0 --no-extra --no-cpu-baseline > gpurun_out/s1_ncu_decode.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_cfg2.csv \
  python bench.py --steps 6 --warmup 4 --repeats 1 --sustain-s 0 --no-extra --no-cpu-baseline > gpurun_out/s1_ncu_launches.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
