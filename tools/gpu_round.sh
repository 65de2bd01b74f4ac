#!/bin/bash
# One GPU round: parity tests, bench, ncu launch list, ncu full capture.  Run via gpurun.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_full.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:ascent_ipm -s 3 -c 1 -o gpurun_out/prof \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
