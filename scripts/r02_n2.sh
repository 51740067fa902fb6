set -x
D=gpurun_out/r02n2; mkdir -p $D
nvidia-smi topo -m > $D/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > $D/bench_n2.json 2> $D/bench_n2.err
timeout 300 python -m pytest tests/test_gpu_pool.py -q > $D/pytest_pool.log 2>&1
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-other > $D/bench_n1.json 2> $D/bench_n1.err
