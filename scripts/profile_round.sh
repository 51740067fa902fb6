#!/bin/bash
# Round profile pass (run on the GPU box through gpurun): bench lines, ncu launch list of one timed step, and ncu
# captures of the dominant kernels.  Every ncu command runs only after the same command exited 0 without ncu.  Reports
# are summarised on the box (scripts/ncu_summary.py) and only the search kernel's .ncu-rep is kept: gpurun_out/ may not
# exceed 64 MiB.  Copy the summaries into profiles/ afterwards.
set -u
O=gpurun_out
TAG=${1:-r01b}
B="python bench.py --pairs 128 --chunk 128 --steps 1 --warmup 1 --no-cpu --no-check --no-e2e"
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis"
summ() { python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt 2>&1; rm -f $O/$1.ncu-rep; }
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_ref.json 2> $O/bench_${TAG}_ref.err
timeout 120 $B > $O/plain_${TAG}.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_${TAG}.csv $B > $O/ncu_launch_${TAG}.log 2>&1
# search kernel: level-0 launch of the timed step (6 search launches in the run: 3 warm-up + 3 timed)
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_tma --launch-skip 5 --launch-count 1 \
  -o $O/prof_${TAG}_search_l0 $B > $O/ncu_search_${TAG}.log 2>&1
python scripts/ncu_summary.py $O/prof_${TAG}_search_l0.ncu-rep > $O/prof_${TAG}_search_l0.txt 2>&1
# regularisation kernels of level 0 in the timed step: the 16x16 stage (first six) and the 2x2 stage (last six)
timeout 600 ncu $LIGHT --clock-control none --kernel-name regex:"k_reg_(eval|fix|classify)" --launch-skip 120 --launch-count 6 \
  -o $O/prof_${TAG}_reg_l0_bs16 $B > $O/ncu_reg16_${TAG}.log 2>&1; summ prof_${TAG}_reg_l0_bs16
timeout 600 ncu $LIGHT --clock-control none --kernel-name regex:"k_reg_(eval|fix|classify)" --launch-skip 138 --launch-count 6 \
  -o $O/prof_${TAG}_reg_l0_bs2 $B > $O/ncu_reg2_${TAG}.log 2>&1; summ prof_${TAG}_reg_l0_bs2
# HBM-bound kernels of the timed step: pad, pyrDown x2 ... last divide, export
timeout 600 ncu $LIGHT --clock-control none --kernel-name regex:"k_(export|pyrdown|pad|divide|copy_mvs)" --launch-skip 15 --launch-count 15 \
  -o $O/prof_${TAG}_hbm $B > $O/ncu_hbm_${TAG}.log 2>&1; summ prof_${TAG}_hbm
timeout 120 python scripts/quarterpel_wrapper.py > $O/quarterpel_${TAG}.json 2> $O/quarterpel_${TAG}.err
timeout 300 ncu $LIGHT --clock-control none --kernel-name regex:"k_(resize_pad|export_subsample)" --launch-skip 2 --launch-count 2 \
  -o $O/prof_${TAG}_wrapper python scripts/quarterpel_wrapper.py > $O/ncu_wrapper_${TAG}.log 2>&1; summ prof_${TAG}_wrapper
timeout 300 python scripts/config_times.py > $O/configs_${TAG}.jsonl 2> $O/configs_${TAG}.err
du -sh $O
