#!/bin/bash
# Round 2: K2v (rollout_cartpole_ws4_kernel) -- parity tests, timing against K2w, role placement sweep, optional ncu capture.
# usage (under gpurun): bash scripts/r2_k2v.sh <tag> [sweep] [ncu]
R=${1:-r2e}
O=gpurun_out
mkdir -p $O
export RL_WS_VARIANT=4
timeout 600 python -m pytest tests/test_gpu_envs.py -x -q -m gpu -k "warp_specialized or set_weights_async" > $O/${R}_pytest_ws.log 2>&1
echo "pytest rc=$?" >> $O/${R}_pytest_ws.log
tail -3 $O/${R}_pytest_ws.log
echo "== K2w (round 1)"; RL_WS_VARIANT=1 timeout 120 python scripts/sweep_rollout.py 1024,4096,8192 160 2>&1 | tail -3
echo "== K2v"; timeout 120 python scripts/sweep_rollout.py 1024,2368,4096,4736,8192 160 2>&1 | tail -5
if [[ "$*" == *sweep* ]]; then
  for roles in 2,5,2,5 2,4,2,4 2,5,3,4 2,4,3,5 1,5,3,4 2,3,2,3 3,5,1,4 2,5,1,4 0,5,2,4 2,1,2,1; do
    echo -n "roles $roles "; RL_WS4_ROLES="$roles" timeout 60 python scripts/sweep_rollout.py 4096 160 2>&1 | tail -1
  done
fi
if [[ "$*" == *ncu* ]]; then
  NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
  timeout 300 $NCU -k "regex:rollout_cartpole_ws4_kernel" -s 2 -c 1 -o $O/${R}_k2v_e4096 python scripts/profile_rollout_small.py 4096 256 160 > $O/${R}_ncu.log 2>&1
  python scripts/ncu_summary.py $O/${R}_k2v_e4096.ncu-rep 16 > $O/${R}_k2v_e4096.md 2>/dev/null
  head -30 $O/${R}_k2v_e4096.md
fi
