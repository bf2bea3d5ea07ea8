#!/bin/bash
# bench.py at N GPUs of one box under torchrun, as the driver launches it:  gpurun --gpus N -- 'bash tools/bench_scaling.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; cat gpurun_out/bench_n$N.json; grep -v "UserWarning\|return func" gpurun_out/bench_n$N.err | tail -5
