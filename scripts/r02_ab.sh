set -x
D=gpurun_out/r02ab; mkdir -p $D
for i in 1 2; do
  timeout 200 python scripts/search_only.py c2 > $D/new_$i.json 2>/dev/null
  BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_oldsearch.so timeout 200 python scripts/search_only.py c2 > $D/old_$i.json 2>/dev/null
done
