#!/bin/bash
# multi-GPU check on one box (run under gpurun --gpus 8): the driver's launch line for N = 8 and 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "N=$N rc=$?"; cat gpurun_out/bench_n$N.json; tail -2 gpurun_out/bench_n$N.err
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "N=1 rc=$?"; cat gpurun_out/bench_n1.json
