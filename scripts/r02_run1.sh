set -x
mkdir -p gpurun_out/r02a
python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
python bench.py > gpurun_out/r02a/bench_default.json 2> gpurun_out/r02a/bench_default.err
for cfg in "128 64 2" "148 148 1" "296 148 2" "296 74 4" "296 296 1" "128 32 4"; do
  set -- $cfg
  python bench.py --pairs $1 --chunk $2 --slots $3 --no-cpu --no-e2e --no-check --steps 4 --warmup 2 > gpurun_out/r02a/bench_p$1_c$2_s$3.json 2>/dev/null
done
nvidia-smi > gpurun_out/r02a/smi.txt; nproc >> gpurun_out/r02a/smi.txt
