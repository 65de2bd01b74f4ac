#!/bin/bash
# One GPU round (run via gpurun, one GPU): parity tests, the bench line, the ncu launch list of the bench, and one
# `ncu --set full` capture each of the thread-per-problem kernel (config 4), the cooperative kernel on a single
# problem and the cooperative kernel on config 5.  Every command runs plain (exit 0) before it runs under ncu.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/benchref.json 2> gpurun_out/benchref.err
if [ "${NCU:-1}" = "1" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_full.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:ascent_ipm -s 3 -c 1 -o gpurun_out/prof_ipm \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  python tools/gpu_single.py coop 1 > gpurun_out/plain_single.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:ascent_coop -s 1 -c 1 -o gpurun_out/prof_coop_b1 \
      python tools/gpu_single.py coop 1 > gpurun_out/ncu_single.log 2>&1
  python tools/gpu_single.py coop 4096 2001 > gpurun_out/plain_cfg5.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:ascent_coop -s 3 -c 1 -o gpurun_out/prof_coop_cfg5 \
      python tools/gpu_single.py coop 4096 2001 > gpurun_out/ncu_cfg5.log 2>&1
fi
ls -la gpurun_out | tail -30
