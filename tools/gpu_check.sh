#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, the ncu launch list and one full capture of the top kernel.
# Usage (from the repo root on the box): bash tools/gpu_check.sh [tests|bench|ncu|all]
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  DD_BENCH_LAYERS=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -40 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/ncu_launch.log 2>&1
  tail -2 gpurun_out/ncu_launch.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 2 -c 3 \
      -o gpurun_out/prof_conv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"post_kp_|compose_rows" -s 6 -c 4 \
      -o gpurun_out/prof_image -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/ncu_full2.log 2>&1
  tail -2 gpurun_out/ncu_full2.log
  ls -la gpurun_out
fi
