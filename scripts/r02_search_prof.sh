set -x
O=gpurun_out/search_r02; mkdir -p $O
summ() { python scripts/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt 2>&1; rm -f $O/$1.ncu-rep; }
for g in c1 c3 c5; do
  timeout 120 python scripts/search_only.py $g > $O/plain_$g.json 2> $O/plain_$g.err || continue
  L=4
  # the finest-level launch of the second call: launches per call = levels; skip = levels + (levels - 1)
  timeout 600 ncu --set full --clock-control none --kernel-name regex:k_search_tma --launch-skip 7 --launch-count 1 -o $O/search_$g python scripts/search_only.py $g > $O/ncu_$g.log 2>&1; summ search_$g
done
