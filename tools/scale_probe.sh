#!/bin/bash
# One multi-GPU call that answers "what limits the host-frame (e2e) path at N GPUs?" (VERDICT r01 item 5).
#   gpurun --gpus 8 -- tools/scale_probe.sh r02
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
{
  echo "== topology =="; nproc; lscpu | grep -i "numa\|socket\|model name"; nvidia-smi topo -m 2>/dev/null | head -20
  echo "== aggregate page-locked host<->device bandwidth (profiles/microbench/pcie_multi.cu) =="
  profiles/microbench/pcie_multi 1.0
} > $OUT/pcie_multi_$TAG.txt 2>&1
N=$(python -c "import torch; print(torch.cuda.device_count())")
# the driver's way: one rank per GPU
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
# one process driving every device through sws_cuda_scale_batch_host()
python tools/batch_host_probe.py > $OUT/batch_host_$TAG.txt 2>&1
cat $OUT/pcie_multi_$TAG.txt; cat $OUT/batch_host_$TAG.txt
python - <<PY
import sys
import json
d = json.loads([l for l in open("$OUT/bench_${TAG}_n$N.json") if l.startswith("{")][-1])
for k in ("n_gpus", "value", "e2e", "e2e_pageable", "e2e_batch_host", "numa"):
    print(k, d.get(k))
PY
