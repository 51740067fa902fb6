set -x
D=gpurun_out/r02q; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
for g in c1 c3 c5 c2; do timeout 200 python scripts/search_only.py $g > $D/search_$g.json 2> $D/search_$g.err; done
timeout 600 python bench.py --no-cpu --steps 10 > $D/bench.json 2> $D/bench.err
