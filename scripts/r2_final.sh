#!/bin/bash
# Round 2 verification pass: every GPU test, compute-sanitizer on the round-2 kernels, the driver's two bench commands,
# an ncu capture of the headline kernel (K2q) and the launch list of a bench run.
# usage (under gpurun): bash scripts/r2_final.sh <tag> [nosan] [noncu]
R=${1:-r2A}
O=gpurun_out
mkdir -p $O
timeout -k 5 1200 python -m pytest tests -x -q -m gpu -s > $O/${R}_pytest_s.log 2>&1
echo "pytest rc=$?"; tail -2 $O/${R}_pytest_s.log
if [[ "$*" != *nosan* ]]; then
  for tool in memcheck racecheck synccheck; do
    timeout -k 5 360 compute-sanitizer --tool $tool python scripts/sanitize_round2.py > $O/${R}_sanitize_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/${R}_sanitize_$tool.log | tail -1
  done
fi
bash scripts/r2_bench.sh $R 1 ref
if [[ "$*" != *noncu* ]]; then
  NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
  timeout -k 5 300 $NCU -k "regex:rollout_cartpole_ws6_kernel" -s 2 -c 1 -o $O/${R}_k2q_e4096 python scripts/profile_rollout_small.py 4096 256 160 > $O/${R}_ncu.log 2>&1
  python scripts/ncu_summary.py $O/${R}_k2q_e4096.ncu-rep 16 > $O/${R}_k2q_e4096.md 2>/dev/null
  head -24 $O/${R}_k2q_e4096.md
  timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 3 > $O/${R}_launches_bench.log 2>&1
  echo "launch list rc=$?"; wc -l $O/${R}_launches.csv
fi
