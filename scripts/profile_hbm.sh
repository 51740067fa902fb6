#!/bin/bash
# ncu --set full of the HBM-bound kernels of one timed step (and of main()'s wrapper), summarised on the box.
set -u
O=gpurun_out; TAG=${1:-r01c}
B="python bench.py --pairs 128 --chunk 128 --steps 1 --warmup 1 --no-cpu --no-check --no-e2e"
summ() { python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt 2>&1; rm -f $O/$1.ncu-rep; }
timeout 120 $B > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --kernel-name regex:"^k_(pad|pyrdown|export)" --launch-skip 4 --launch-count 4 -o $O/prof_${TAG}_hbm_a $B > $O/ncu_hbm_a.log 2>&1; summ prof_${TAG}_hbm_a
timeout 400 ncu --set full --clock-control none --kernel-name regex:"k_(divide2|copy_mvs)" --launch-skip 20 --launch-count 2 -o $O/prof_${TAG}_hbm_b $B > $O/ncu_hbm_b.log 2>&1; summ prof_${TAG}_hbm_b
timeout 400 ncu --set full --clock-control none --kernel-name regex:"k_reg_classify4" --launch-skip 42 --launch-count 1 -o $O/prof_${TAG}_hbm_c $B > $O/ncu_hbm_c.log 2>&1; summ prof_${TAG}_hbm_c
timeout 120 python scripts/quarterpel_wrapper.py > $O/quarterpel_${TAG}.json 2>/dev/null
timeout 300 ncu --set full --clock-control none --kernel-name regex:"k_(resize_pad|export_subsample)" --launch-skip 2 --launch-count 2 -o $O/prof_${TAG}_wrapper python scripts/quarterpel_wrapper.py > $O/ncu_wrapper.log 2>&1; summ prof_${TAG}_wrapper
cat $O/prof_${TAG}_hbm_a.txt $O/prof_${TAG}_hbm_b.txt $O/prof_${TAG}_hbm_c.txt $O/prof_${TAG}_wrapper.txt > $O/prof_${TAG}_hbm_all.txt
du -sh $O
