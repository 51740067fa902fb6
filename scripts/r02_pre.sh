# Byte-shifted copies (PRE search kernels) against the funnel-shift kernels of the same library (BBME_SEARCH_PRE=0), same box.
set -x
D=gpurun_out/r02pre; mkdir -p $D
timeout 1200 python -m pytest tests -x -q -m gpu > $D/pytest_gpu.log 2>&1
for i in 1 2; do
  for g in c2 c1 c3; do
    timeout 200 python scripts/search_only.py $g > $D/pre_${g}_$i.json 2>$D/pre_${g}_$i.err
    BBME_SEARCH_PRE=0 timeout 200 python scripts/search_only.py $g > $D/shf_${g}_$i.json 2>/dev/null
  done
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-other > $D/bench.json 2> $D/bench.err
