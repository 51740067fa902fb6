set -x
D=gpurun_out/r02i; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128.json 2> $D/prof128.err
timeout 300 python scripts/reg_profile.py 1 8 > $D/prof1.json 2> $D/prof1.err
for cfg in "32 4" "48 3" "64 3" "128 2"; do set -- $cfg; timeout 300 python bench.py --no-cpu --no-other --no-check --e2e-chunk $1 --e2e-slots $2 --steps 4 > $D/bench_e2e_c$1_s$2.json 2>$D/bench_e2e_c$1_s$2.err; done
