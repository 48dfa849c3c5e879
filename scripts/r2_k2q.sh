#!/bin/bash
# Round 2: K2q (rollout_cartpole_ws6_kernel) -- timing against K2w, parity tests, role placements, per-phase clocks.
# usage (under gpurun): bash scripts/r2_k2q.sh <tag> [roles] [clocks]
R=${1:-r2w}
O=gpurun_out
mkdir -p $O
T="timeout -k 5"
echo "== K2w"; RL_WS_VARIANT=1 $T 60 python scripts/sweep_rollout.py 4096 160 2>&1 | tail -1
echo "== K2q"; RL_WS_VARIANT=6 $T 60 python scripts/sweep_rollout.py 1024,4096,4736 160 2>&1 | tail -3
RL_WS_VARIANT=6 $T 300 python -m pytest tests/test_gpu_envs.py tests/test_gpu_fullsize.py -x -q -m gpu -k "warp_specialized or set_weights_async or bench_size" > $O/${R}_pytest_ws6.log 2>&1
echo "variant 6 pytest rc=$?" | tee -a $O/${R}_pytest_ws6.log
tail -3 $O/${R}_pytest_ws6.log
if [[ "$*" == *roles* ]]; then
  # role tables (RL_WS6_TABLE: warp w -> policy index | 8 dynamics | 9 aux | 15 idle; warp w runs on sub-partition w % 4)
  for table in 0,1,2,8,3,4,5,15,6,15,9,15 0,1,2,8,3,4,5,15,6,7,9,15 0,2,4,8,1,3,5,6,15,15,9,15 0,1,2,8,3,4,5,15,6,9,15,15 0,1,2,8,3,4,5,9,6,15,15,15; do
    echo -n "K2q table $table "; RL_WS_VARIANT=6 RL_WS6_TABLE="$table" $T 40 python scripts/sweep_rollout.py 4096 160 2>&1 | tail -1
  done
fi
if [[ "$*" == *clocks* ]]; then $T 120 bash scripts/ws_clocks.sh $R "6"; fi
