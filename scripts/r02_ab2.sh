# Same-box A/B of two builds of the search kernel over the BASELINE geometries: the in-tree library against
# blockbasedmotionestimation_b200/libbbme_oldsearch.so (selected through BBME_LIB), interleaved, two rounds.
set -x
D=gpurun_out/r02ab2; mkdir -p $D
for i in 1 2; do
  for g in c2 c1 c3 c5; do
    timeout 200 python scripts/search_only.py $g > $D/new_${g}_$i.json 2>/dev/null
    BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_oldsearch.so timeout 200 python scripts/search_only.py $g > $D/old_${g}_$i.json 2>/dev/null
  done
done
