set -x
D=gpurun_out/r02n4; mkdir -p $D
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --no-other --no-check"
timeout 600 $T > $D/bench_default.json 2> $D/bench_default.err
BBME_HOST_THREADS=4 timeout 600 $T > $D/bench_t4.json 2> $D/bench_t4.err
BBME_HOST_THREADS=12 timeout 600 $T > $D/bench_t12.json 2> $D/bench_t12.err
