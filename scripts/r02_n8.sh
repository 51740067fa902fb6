set -x
D=gpurun_out/r02n8; mkdir -p $D
timeout 600 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
nvidia-smi topo -m > $D/topo.txt 2>&1; nproc >> $D/topo.txt; free -g >> $D/topo.txt
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > $D/bench_n$N.json 2> $D/bench_n$N.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > $D/bench_n1.json 2> $D/bench_n1.err
timeout 300 python - > $D/pool.json 2> $D/pool.err <<'PY'
import json, time, sys
sys.path.insert(0, '.')
import numpy as np
import blockbasedmotionestimation_b200 as bb
from blockbasedmotionestimation_b200.synth import make_pair, seed_for
W, H, SS, BS = 1920, 1080, [80] * 3, [16] * 3
pairs = [make_pair(H, W, seed_for(4, i), shift=(5 - (i % 11), (i % 7) - 3), patches=12, max_patch_shift=40) for i in range(16)]
n = 512
f1 = [pairs[i % 16][0] for i in range(n)]; f2 = [pairs[i % 16][1] for i in range(n)]
with bb.Pool(W, H, SS, BS, chunk_pairs=32, slots=2) as pool:
    pool.estimate_batch(f1[:64], f2[:64])
    t = time.perf_counter(); out = pool.estimate_batch(f1, f2); dt = time.perf_counter() - t
    same = all(np.array_equal(out[i], out[i % 16]) for i in range(n))
    print(json.dumps({"what": "bbme_pool_estimate_batch: one process, one host thread per GPU, pageable numpy buffers", "devices": pool.device_count,
                      "pairs": n, "pairs_per_s": n / dt, "identical_across_devices": bool(same)}))
PY
