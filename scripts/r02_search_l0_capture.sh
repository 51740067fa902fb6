# ncu --set full of the level-0 search launch of a 128-pair chunk (the capture behind profiles/r02_search_l0_ncu.txt), after the
# same command has exited 0 without ncu.  Summary and dynamic instruction mix by scripts/ncu_summary.py on the box.
set -u
O=gpurun_out/prof_r02c; mkdir -p $O
B="python bench.py --pairs 128 --chunk 128 --slots 1 --steps 1 --warmup 1 --no-cpu --no-check --no-e2e --no-other"
timeout 120 $B > $O/plain.log 2>&1 || exit 1
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_tma --launch-skip 5 --launch-count 1 -o $O/search_l0 $B > $O/ncu_search.log 2>&1
python scripts/ncu_summary.py $O/search_l0.ncu-rep > $O/search_l0.txt 2>&1; rm -f $O/search_l0.ncu-rep
