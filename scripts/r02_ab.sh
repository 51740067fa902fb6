# Same-box A/B of the search kernel on config 2 (128 x 1080p pairs): the in-tree library against a library linked with the
# search_tma.cu of commit a26cbc3 (before the round-2 search changes), built with
#   git show a26cbc3:blockbasedmotionestimation_b200/csrc/search_tma.cu > blockbasedmotionestimation_b200/csrc/_search_old.cu
#   nvcc <Makefile flags> -c .../_search_old.cu -o /tmp/search_old.o ; nvcc -shared -o .../libbbme_oldsearch.so <other .o> /tmp/search_old.o
# and selected through BBME_LIB.  Results quoted in DESIGN.md 3.1: 26.90 vs 26.89 ms for the final build; 27.09 / 27.20 vs 26.91 /
# 26.98 ms for two intermediate versions (run-time ring depth; ring slots cached in two extra live registers).
set -x
D=gpurun_out/r02ab; mkdir -p $D
for i in 1 2; do
  timeout 200 python scripts/search_only.py c2 > $D/new_$i.json 2>/dev/null
  BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_oldsearch.so timeout 200 python scripts/search_only.py c2 > $D/old_$i.json 2>/dev/null
done
