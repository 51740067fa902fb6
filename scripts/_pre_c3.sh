D=gpurun_out/r02pre; mkdir -p $D
BBME_SEARCH_PRE_FORCE=1 timeout 200 python scripts/search_only.py c3 > $D/preforce_c3.json 2>$D/preforce_c3.err
timeout 200 python scripts/search_only.py c3 > $D/shf_c3_3.json 2>/dev/null
BBME_SEARCH_PRE_FORCE=1 timeout 300 python -m pytest tests/test_gpu_big_golden.py -q -k "c3" > $D/pytest_c3_force.log 2>&1
