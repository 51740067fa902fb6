set -u
O=gpurun_out; TAG=r01c
B="python bench.py --pairs 128 --chunk 128 --steps 1 --warmup 1 --no-cpu --no-check --no-e2e"
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis"
summ() { python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt 2>&1; rm -f $O/$1.ncu-rep; }
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_ref.json 2> $O/bench_${TAG}_ref.err
timeout 120 $B > $O/plain_${TAG}.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_${TAG}.csv $B > $O/ncu_launch_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name regex:k_search_tma --launch-skip 5 --launch-count 1 -o $O/prof_${TAG}_search_l0 $B > $O/ncu_search_${TAG}.log 2>&1; summ prof_${TAG}_search_l0
timeout 600 ncu --set full --clock-control none --kernel-name regex:"k_reg_(eval|fix|classify)" --launch-skip 120 --launch-count 3 -o $O/prof_${TAG}_reg_l0_bs16 $B > $O/ncu_reg16_${TAG}.log 2>&1; summ prof_${TAG}_reg_l0_bs16
timeout 600 ncu --set full --clock-control none --kernel-name regex:"k_(export|divide2)" --launch-skip 17 --launch-count 2 -o $O/prof_${TAG}_hbm $B > $O/ncu_hbm_${TAG}.log 2>&1; summ prof_${TAG}_hbm
du -sh $O
