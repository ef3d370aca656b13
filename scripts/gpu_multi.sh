#!/bin/bash
# 2-GPU check: bench under torchrun (weak scaling, one model per rank) and the replicate work queue.
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m pytest tests/test_gpu_tp.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/tp_test_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 260 --warmup 26 > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err; echo "exit $?" >> gpurun_out/bench_n$N.err
mkdir -p /tmp/bs && (time timeout 900 python -m locator_b200 --vcf tests/golden/data/test_genotypes.vcf.gz --sample_data tests/golden/data/test_sample_data.txt --out /tmp/bs/run --seed 12345 --keras_verbose 0 --bootstrap --nboots 8 --max_epochs 60 --gpus $N) > gpurun_out/bootstrap_n$N.log 2>&1
ls /tmp/bs | wc -l >> gpurun_out/bootstrap_n$N.log
(time timeout 900 python -m locator_b200 --vcf tests/golden/data/test_genotypes.vcf.gz --sample_data tests/golden/data/test_sample_data.txt --out /tmp/bs/run1 --seed 12345 --keras_verbose 0 --bootstrap --nboots 8 --max_epochs 60 --gpus 1) > gpurun_out/bootstrap_n1.log 2>&1
for b in 0 3 7; do cmp /tmp/bs/run_boot${b}_predlocs.txt /tmp/bs/run1_boot${b}_predlocs.txt && echo "boot $b identical across --gpus" >> gpurun_out/bootstrap_n$N.log; done
tail -3 gpurun_out/tp_test_n$N.log; cut -c1-500 gpurun_out/bench_n$N.log; python -c "import json;d=[json.loads(l) for l in open('gpurun_out/bench_n$N.log') if l.startswith('{')][0];print('TP', d.get('tensor_parallel'));print('value',d['value'],'e2e',d['e2e']['value'])"; tail -3 gpurun_out/bench_n$N.err; tail -8 gpurun_out/bootstrap_n$N.log; tail -4 gpurun_out/bootstrap_n1.log
