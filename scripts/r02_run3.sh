set -x
D=gpurun_out/r02c; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
timeout 600 python bench.py --no-cpu > $D/bench_default.json 2> $D/bench_default.err
timeout 120 python scripts/single_pair.py > $D/single_pair_fused.json 2>&1
for cs in 1 4; do BBME_REG_CLUSTER=$cs timeout 120 python scripts/single_pair.py > $D/single_pair_cs$cs.json 2>&1; done
