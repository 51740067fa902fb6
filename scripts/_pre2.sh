D=gpurun_out/r02pre2; mkdir -p $D
timeout 1200 python -m pytest tests -x -q -m gpu > $D/pytest_gpu.log 2>&1
for i in 1 2; do
  for g in c2 c1 c3 c5; do
    timeout 200 python scripts/search_only.py $g > $D/new_${g}_$i.json 2>$D/new_${g}_$i.err
  done
done
