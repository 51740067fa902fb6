set -x
D=gpurun_out/r02p; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
for g in c1 c3 c5 c2; do timeout 200 python scripts/search_only.py $g > $D/search_$g.json 2> $D/search_$g.err; done
for st in 5 8; do for g in c1 c2; do BBME_SEARCH_STAGES=$st timeout 200 python scripts/search_only.py $g > $D/search_${g}_st$st.json 2> /dev/null; done; done
timeout 600 python bench.py --no-cpu --steps 10 > $D/bench.json 2> $D/bench.err
