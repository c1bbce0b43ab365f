#!/bin/bash
# The driver's scaling run on ONE 8-GPU box: python bench.py at N = 1, then torchrun at N = 2, 4, 8, back to back.
mkdir -p gpurun_out; : > gpurun_out/scale_1_2_4_8_r02.jsonl
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 >> gpurun_out/scale_1_2_4_8_r02.jsonl 2> gpurun_out/scale_r02.err; echo "n1 rc=$?"
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 >> gpurun_out/scale_1_2_4_8_r02.jsonl 2>> gpurun_out/scale_r02.err; echo "n$n rc=$?"
done
python - <<PY
import json
for ln in open("gpurun_out/scale_1_2_4_8_r02.jsonl"):
    try: d = json.loads(ln)
    except Exception: continue
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["value"], d["e2e"].get("h2d_gbs_per_gpu"), d["ref5"]["ms_per_step"], d["ref5"]["collective_ms_last"], d["ref5"].get("rel_l2_vs_oracle"))
PY
