#!/bin/bash
# Multi-GPU round trip (gpurun --gpus N): 2-device parity test, then bench.py under torchrun exactly as the driver
# launches it, for every N in the argument list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.log 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_host_cpp.py -m gpu -q --timeout=200 --timeout-method=thread > gpurun_out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
port=29510
for n in "$@"; do
  port=$((port+1))
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  fi
  echo "bench n=$n rc=$?" >> gpurun_out/pytest_multi.log
done
tail -4 gpurun_out/pytest_multi.log
