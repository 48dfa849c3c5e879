#!/bin/bash
# Round 2: K2z (rollout_cartpole_ws5_kernel) -- parity tests, timing against K2w / K2v, per-phase clocks.
# usage (under gpurun): bash scripts/r2_k2z.sh <tag> [roles]
R=${1:-r2v}
O=gpurun_out
mkdir -p $O
RL_WS_VARIANT=5 timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_fullsize.py -x -q -m gpu -k "warp_specialized or set_weights_async or bench_size" > $O/${R}_pytest_ws.log 2>&1
echo "pytest rc=$?" >> $O/${R}_pytest_ws.log
tail -4 $O/${R}_pytest_ws.log
echo "== K2w"; RL_WS_VARIANT=1 timeout 120 python scripts/sweep_rollout.py 1024,4096,8192 160 2>&1 | tail -3
echo "== K2v"; RL_WS_VARIANT=4 timeout 120 python scripts/sweep_rollout.py 1024,4096,8192 160 2>&1 | tail -3
echo "== K2z"; RL_WS_VARIANT=5 timeout 120 python scripts/sweep_rollout.py 1024,2368,4096,4736,8192 160 2>&1 | tail -5
if [[ "$*" == *roles* ]]; then
  for roles in 2,5,2,5 2,4,2,4 2,5,3,4 2,4,3,5 1,5,3,4 2,3,2,3 3,5,1,4 2,5,1,4 0,5,2,4 2,1,2,1 0,4,1,5 0,5,1,4 3,4,2,5; do
    echo -n "roles $roles "; RL_WS_VARIANT=5 RL_WS4_ROLES="$roles" timeout 60 python scripts/sweep_rollout.py 4096 160 2>&1 | tail -1
  done
fi
bash scripts/ws_clocks.sh $R 5
