set -x
D=gpurun_out/r02n2b; mkdir -p $D
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
timeout 600 $T --steps 10 --warmup 3 --no-other --no-e2e > $D/bench_n2_gather.json 2> $D/bench_n2_gather.err
timeout 600 $T --steps 10 --warmup 3 --no-other --no-e2e --no-gather > $D/bench_n2_nogather.json 2> $D/bench_n2_nogather.err
timeout 600 $T --steps 10 --warmup 3 > $D/bench_n2_full.json 2> $D/bench_n2_full.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-e2e > $D/bench_n1.json 2> $D/bench_n1.err
