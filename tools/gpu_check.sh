#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and one full capture of the GEMM kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
  tail -25 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_ref.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  tail -2 gpurun_out/ncu_launches.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 120 -c 4 -f -o gpurun_out/prof_gemm \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
  tail -2 gpurun_out/ncu_gemm.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 10 -c 1 -f -o gpurun_out/prof_attn \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
  tail -2 gpurun_out/ncu_attn.log
fi
ls -la gpurun_out | tail -20
