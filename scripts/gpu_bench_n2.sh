#!/bin/bash
# bench.py under torchrun on 2 GPUs, as the driver launches it (short run).
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 260 --warmup 26 --cpu-steps 2 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "exit: $?"; grep '^{' gpurun_out/n2_bench.json | cut -c1-1500; tail -3 gpurun_out/n2_bench.err
