set -x
D=gpurun_out/r02g; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128_t1024.json 2> $D/prof128_t1024.err
BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_t512.so BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128_t512.json 2> $D/prof128_t512.err
timeout 300 python scripts/reg_profile.py 1 8 > $D/prof1_t1024.json 2> $D/prof1.err
BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_t512.so timeout 300 python scripts/reg_profile.py 1 8 > $D/prof1_t512.json 2>> $D/prof1.err
timeout 300 python scripts/reg_profile.py 16 8 > $D/prof16_t1024.json 2> $D/prof16.err
BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_t512.so timeout 300 python scripts/reg_profile.py 16 8 > $D/prof16_t512.json 2>> $D/prof16.err
