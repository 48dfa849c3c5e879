#!/bin/bash
# Round profile pass (run under gpurun, one GPU): launch list of a bench run + one --set full capture per hot kernel.
# Outputs go to gpurun_out/; summaries are made from them with scripts/ncu_summary.py and committed under profiles/.
R=${1:-r1}
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 3 > $O/${R}_bench_under_ncu.log 2>&1
$NCU -k "regex:rollout_cartpole_group_kernel" -s 2 -c 1 -o $O/${R}_k2c_e4096 python scripts/profile_rollout_small.py 4096 256 0 > $O/ncu.log 2>&1
$NCU -k "regex:rollout_cartpole_group_kernel" -s 2 -c 1 -o $O/${R}_k2c_e1m python scripts/profile_rollout_small.py 1048576 32 1 >> $O/ncu.log 2>&1
$NCU -k "regex:rollout_cartpole_tc_kernel" -s 2 -c 1 -o $O/${R}_k2t_e1m python scripts/profile_rollout_small.py 1048576 32 128 >> $O/ncu.log 2>&1
$NCU -k "regex:rollout_cartpole_tc_kernel" -s 2 -c 1 -o $O/${R}_k2t_e16k python scripts/profile_rollout_small.py 16384 256 128 >> $O/ncu.log 2>&1
$NCU -k "regex:mlp_pass_tc_kernel<\(int\)1, \(int\)4>" -s 2 -c 1 -o $O/${R}_k6_value_tc python scripts/profile_kernels.py update >> $O/ncu.log 2>&1
$NCU -k "regex:mlp_pass_tc_kernel<\(int\)2, \(int\)3>" -s 2 -c 1 -o $O/${R}_k5_fvp_tc python scripts/profile_kernels.py update >> $O/ncu.log 2>&1
RL_PASS_KERNEL=ffma $NCU -k "regex:mlp_pass_kernel<\(int\)5, \(int\)1" -s 1 -c 1 -o $O/${R}_k6_value python scripts/profile_kernels.py update >> $O/ncu.log 2>&1
RL_PASS_KERNEL=ffma $NCU -k "regex:mlp_pass_kernel<\(int\)5, \(int\)2, \(int\)4, \(int\)3" -s 1 -c 1 -o $O/${R}_k5_fvp python scripts/profile_kernels.py update >> $O/ncu.log 2>&1
$NCU -k "regex:env_step_kernel" -s 2 -c 1 -o $O/${R}_k1_step python scripts/profile_kernels.py step >> $O/ncu.log 2>&1
$NCU -k "regex:gae_scan_kernel|value_forward_pairs_kernel" -s 1 -c 3 -o $O/${R}_k3_scan python scripts/profile_kernels.py scan >> $O/ncu.log 2>&1
$NCU -k "regex:replay_copy_kernel<float, \(int\)8>|sample_gather_kernel|sample_scan_kernel|replay_book_kernel" -s 4 -c 4 -o $O/${R}_k4_replay python scripts/profile_extra.py dqn >> $O/ncu.log 2>&1
$NCU -k "regex:rollout_seq_kernel" -s 1 -c 1 -o $O/${R}_k8_gru python scripts/profile_extra.py gru >> $O/ncu.log 2>&1
# gpurun copies back at most 64 MiB: summarise every report here, keep only the reports of the two tensor-core kernels
for rep in $O/${R}_*.ncu-rep; do
  python scripts/ncu_summary.py $rep 16 > ${rep%.ncu-rep}.md 2>/dev/null
done
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
for rep in $O/${R}_*.ncu-rep; do
  case $rep in *_tc.ncu-rep) ;; *) rm -f $rep ;; esac
done
tail -3 $O/ncu.log
