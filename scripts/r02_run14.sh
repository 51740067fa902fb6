set -x
D=gpurun_out/r02o; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128_t640.json 2> $D/prof128_t640.err
for T in 512 768; do BBME_LIB=$PWD/blockbasedmotionestimation_b200/libbbme_t$T.so timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128_t$T.json 2> $D/prof128_t$T.err; done
timeout 300 python scripts/reg_profile.py 1 8 > $D/prof1.json 2> $D/prof1.err
timeout 300 python scripts/reg_profile.py 16 8 > $D/prof16.json 2> $D/prof16.err
