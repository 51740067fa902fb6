set -x
D=gpurun_out/r02l; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128.json 2> $D/prof128.err
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 1 8 > $D/prof1.json 2> $D/prof1.err
timeout 300 python scripts/reg_profile.py 16 8 > $D/prof16.json 2> $D/prof16.err
