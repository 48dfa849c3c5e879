#!/bin/bash
# Round 2: the driver's bench commands at N GPUs (+ the multi-GPU parity script), outputs under gpurun_out/<tag>_*.
# usage (under gpurun [--gpus N]): bash scripts/r2_bench.sh <tag> <N> [ref] [sweep]
R=${1:-r2g}; N=${2:-1}
O=gpurun_out
mkdir -p $O
if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
if [[ "$*" == *ref* ]]; then
  timeout 600 $LAUNCH bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $O/${R}_bench_ref_$N.json 2> $O/${R}_bench_ref_$N.err
  echo "reference rc=$?"; cut -c1-600 $O/${R}_bench_ref_$N.json
fi
timeout 1500 $LAUNCH bench.py --gpus $N > $O/${R}_bench_$N.json 2> $O/${R}_bench_$N.err
echo "bench rc=$?"; tail -3 $O/${R}_bench_$N.err; cut -c1-1500 $O/${R}_bench_$N.json
if [ "$N" != "1" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/multigpu_check.py > $O/${R}_multigpu_check_$N.txt 2>&1
  echo "multigpu_check rc=$?"; grep -v "^W\|^\[W\|NCCL" $O/${R}_multigpu_check_$N.txt | tail -5
fi
if [[ "$*" == *sweep* ]]; then
  timeout 900 $LAUNCH bench.py --sweep --gpus $N > $O/${R}_sweep_$N.jsonl 2> $O/${R}_sweep_$N.err
  echo "sweep rc=$?"; cat $O/${R}_sweep_$N.jsonl | cut -c1-400
fi
