#!/bin/bash
# Round 2: the fused tail of the tensor-core passes (pass_tail.cuh) -- parity tests at 1 and 2 GPUs, update timing with the
# tail off / on.  usage (under gpurun --gpus 2): bash scripts/r2_tail.sh <tag>
R=${1:-r2h}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_dqn.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py tests/test_gpu_train.py -x -q -m gpu > $O/${R}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/${R}_pytest.log
for tail in 0 1; do
  echo "== RL_PASS_TAIL=$tail, 1 GPU"
  RL_PASS_TAIL=$tail CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/time_update.py 2>&1 | tail -6
done
for tail in 0 1; do
  echo "== RL_PASS_TAIL=$tail, 2 GPUs"
  RL_PASS_TAIL=$tail timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --quick --steps 20 > $O/${R}_bench2_tail$tail.json 2> $O/${R}_bench2_tail$tail.err
  python - <<PY
import json
d=json.load(open("$O/${R}_bench2_tail$tail.json"))
u=d["update"]; print({k:u[k] for k in ("adv_est_ms","trpo_policy_ms","critic_80_adam_ms","total_ms","all_reduce")})
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/multigpu_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL\|OMP\|\*\*\*" | tail -4
