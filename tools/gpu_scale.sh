#!/bin/bash
# Strong-scaling curve of the bench (config 4, one 65,536 batch index-sharded over N GPUs) and config 5 on N GPUs.
# Run under `gpurun --gpus 8`; lines go to gpurun_out/r02_scale_*.json.
set -u
mkdir -p gpurun_out
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scale_cfg4_n$N.json 2> gpurun_out/r02_scale_cfg4_n$N.err
  tail -c 400 gpurun_out/r02_scale_cfg4_n$N.json; echo
done
for N in 1 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --config 5 --steps 3 --no-cpu-baseline > gpurun_out/r02_scale_cfg5_n1.json 2> gpurun_out/r02_scale_cfg5_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) \
      bench.py --gpus $N --config 5 --steps 3 --no-weak > gpurun_out/r02_scale_cfg5_n$N.json 2> gpurun_out/r02_scale_cfg5_n$N.err
  fi
  tail -c 300 gpurun_out/r02_scale_cfg5_n$N.json; echo
done
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -m gpu -q -k "nccl or device_list" > gpurun_out/r02_pytest_multi_8gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_multi_8gpu.log
