#!/bin/bash
# BASELINE config 3 (608x608, batch 256 sharded by image) on 2 / 4 / 8 GPUs of one box, plus the default 416^2 weak-scaling
# line on 8 GPUs.  Run under `gpurun --gpus 8`; the 1-GPU lines come from a 1-GPU box (tools/scale_608.sh 1).
# usage: bash tools/scale_608.sh <tag> [ngpus...]
TAG=${1:-r2}; shift
NS=${@:-"2 4 8"}
mkdir -p gpurun_out
COMMON="--steps 40 --warmup 5 --settle-seconds 1.5 --no-cpu-baseline"
port=29510
for n in $NS; do
  port=$((port+1))
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --size 608 --global-batch 256 $COMMON > gpurun_out/${TAG}_608_n1.json 2> gpurun_out/${TAG}_608_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --size 608 --global-batch 256 $COMMON > gpurun_out/${TAG}_608_n$n.json 2> gpurun_out/${TAG}_608_n$n.err
  fi
  echo "608 gb256 N=$n exit $?"; tail -1 gpurun_out/${TAG}_608_n$n.json | cut -c1-260
done
for n in $NS; do
  if [ "$n" = "8" ] || [ "$n" = "1" ]; then
    port=$((port+1))
    if [ "$n" = "1" ]; then
      timeout 300 python bench.py --gpus 1 $COMMON > gpurun_out/${TAG}_416_n1.json 2> gpurun_out/${TAG}_416_n1.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $COMMON > gpurun_out/${TAG}_416_n$n.json 2> gpurun_out/${TAG}_416_n$n.err
    fi
    echo "416 b64/gpu N=$n exit $?"; tail -1 gpurun_out/${TAG}_416_n$n.json | cut -c1-260
  fi
done
