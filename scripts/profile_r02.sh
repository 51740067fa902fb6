#!/bin/bash
# Round-2 profile pass (run on the GPU box through gpurun): bench lines, ncu launch list of one timed step, ncu --set full
# captures of the dominant kernels (each only after the same command exited 0 without ncu), in-kernel phase profile of the
# regularisation.  Reports are summarised on the box (scripts/ncu_summary.py); gpurun_out/ may not exceed 64 MiB.
set -u
O=gpurun_out/prof_r02; mkdir -p $O
B="python bench.py --pairs 128 --chunk 128 --slots 1 --steps 1 --warmup 1 --no-cpu --no-check --no-e2e --no-other"
summ() { python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt 2>&1; rm -f $O/$1.ncu-rep; }
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 120 $B > $O/plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv $B > $O/ncu_launch.log 2>&1
# search kernel: level-0 launch of the timed step (6 search launches in the run: 3 warm-up + 3 timed)
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_tma --launch-skip 5 --launch-count 1 -o $O/search_l0 $B > $O/ncu_search.log 2>&1; summ search_l0
# regularisation: the level-1 and level-0 launches of the timed step (6 launches in the run)
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_reg_level --launch-skip 4 --launch-count 2 -o $O/reg_level $B > $O/ncu_reg.log 2>&1; summ reg_level
# HBM-bound kernels of the timed step
timeout 600 ncu --set full --clock-control none --kernel-name regex:"k_(export|pyrdown|pad|copy_mvs|shift4)" --launch-skip 9 --launch-count 9 -o $O/hbm $B > $O/ncu_hbm.log 2>&1; summ hbm
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 128 1 > $O/regprof128.json 2> $O/regprof128.err
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 1 8 > $O/regprof1.json 2> $O/regprof1.err
timeout 120 python scripts/quarterpel_wrapper.py > $O/quarterpel.json 2> $O/quarterpel.err
du -sh $O
