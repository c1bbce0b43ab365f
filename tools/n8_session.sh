#!/bin/bash
# 8-GPU session: the driver's N=8 bench launch, the multi-GPU pytest (NCCL), the host-fabric probe
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu_r02.txt 2>&1; lscpu | grep -E "NUMA|^CPU\(s\)|Model name" >> gpurun_out/topo_8gpu_r02.txt; free -g >> gpurun_out/topo_8gpu_r02.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r02.json 2> gpurun_out/bench_n8_r02.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/bench_n8_r02.err; cut -c1-400 gpurun_out/bench_n8_r02.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/h2d_probe.py > gpurun_out/h2d_probe_r02.jsonl 2> gpurun_out/h2d_probe.err; echo "h2d rc=$?"; cat gpurun_out/h2d_probe_r02.jsonl
timeout 600 python -m pytest tests/test_multigpu_nccl.py tests/test_sharded_gpu.py -q -m gpu -x -p no:cacheprovider -rs -s -k "nccl or every_device" > gpurun_out/multigpu_n8box_r02.log 2>&1; echo "pytest rc=$?"; grep -E "world=|passed|failed" gpurun_out/multigpu_n8box_r02.log | tail -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n4_r02.json 2> gpurun_out/bench_n4_r02.err; echo "bench n4 rc=$?"; cut -c1-300 gpurun_out/bench_n4_r02.json
