set -x
D=gpurun_out/r02d; mkdir -p $D
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 128 1 > $D/prof128_cs1.json 2> $D/prof128_cs1.err
BBME_REG_PROFILE=1 timeout 300 python scripts/reg_profile.py 1 1 8 > $D/prof1.json 2> $D/prof1.err
timeout 300 python scripts/reg_profile.py 128 2 4 8 > $D/cluster128.json 2> $D/cluster128.err
timeout 300 python scripts/reg_profile.py 16 1 2 4 8 > $D/cluster16.json 2> $D/cluster16.err
