D=gpurun_out/r02sweep2; mkdir -p $D
for cfg in "16 8 64 4" "32 6 48 4" "32 8 64 3" "22 6 43 3" "32 4 64 3"; do
  set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --no-other --no-cpu --no-check --chunk $1 --slots $2 --e2e-chunk $3 --e2e-slots $4 > $D/b_$1_$2_$3_$4.json 2> $D/err_$1_$2.log
done
