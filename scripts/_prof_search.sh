set -u
O=gpurun_out/prof_r02b; mkdir -p $O
B="python bench.py --pairs 128 --chunk 128 --steps 1 --warmup 1 --no-cpu --no-check --no-e2e --no-other"
timeout 120 $B > $O/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_tma --launch-skip 5 --launch-count 1 -o $O/search_l0 $B > $O/ncu_search.log 2>&1
python scripts/ncu_summary.py $O/search_l0.ncu-rep > $O/search_l0.txt 2>&1
ncu -i $O/search_l0.ncu-rep --page source --csv > $O/search_l0_source.csv 2>/dev/null
gzip -f $O/search_l0_source.csv; rm -f $O/search_l0.ncu-rep
