set -x
mkdir -p gpurun_out/r02b
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b/pytest.log
timeout 600 python bench.py > gpurun_out/r02b/bench_default.json 2> gpurun_out/r02b/bench_default.err
BBME_REG_LEGACY=1 timeout 300 python bench.py --no-cpu --no-other --no-e2e > gpurun_out/r02b/bench_legacy.json 2> gpurun_out/r02b/bench_legacy.err
timeout 120 python scripts/single_pair.py > gpurun_out/r02b/single_pair_fused.json 2>&1
BBME_REG_LEGACY=1 timeout 120 python scripts/single_pair.py > gpurun_out/r02b/single_pair_legacy.json 2>&1
for cs in 1 2 4; do BBME_REG_CLUSTER=$cs timeout 120 python scripts/single_pair.py > gpurun_out/r02b/single_pair_cs$cs.json 2>&1; done
