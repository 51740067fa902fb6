# Final round-2 build on four GPUs of one box: the torchrun bench line only (value, gather, e2e against the host ceiling).
D=gpurun_out/r02final4; mkdir -p $D
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > $D/bench_n4.json 2> $D/bench_n4.err
