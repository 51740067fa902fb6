"""The N>1 host logic on CPU: two gloo ranks shard a batch of pairs, each 'processes' its shard, results are gathered
(the only collective on the path).  The per-pair computation is stood in by the oracle on tiny frames."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blockbasedmotionestimation_b200.shard import ResultGather, gather_fields, my_shard, shard_bounds
from blockbasedmotionestimation_b200.synth import make_pair


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1024, 1023):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    a, b = my_shard(n_pairs)
    local = []
    for p in range(a, b):
        f1, f2 = make_pair(32, 48, 700 + p, shift=(p % 3 - 1, 1), patches=0, noise=0)
        flow, _ = ob.estimate(f1, f2, [8], [4], 2)
        local.append(torch.from_numpy(flow[::2, ::2].astype(np.int16)))  # compact 2x2-granular field
    local = torch.stack(local) if local else torch.zeros((0, 16, 24, 2), dtype=torch.int16)
    full = gather_fields(local, n_pairs)             # all ranks
    on0 = gather_fields(local, n_pairs, dst=0)       # rank 0 only
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)         # the bench's max-over-ranks timing reduction
    assert float(t.item()) == float(world)
    assert (on0 is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(on0, full)
    # the bench's per-step gather on rank 0 (on CPU tensors: the dist.gather path; on GPUs it maps rank 0's buffer over IPC)
    rg = ResultGather(local, n_pairs, n_buffers=2)
    assert not rg.ipc and "gather" in rg.how
    for which in (0, 1):
        rg.push(local, which)
    dist.barrier()
    if rank == 0:
        assert torch.equal(rg.result(0), full) and torch.equal(rg.result(1), full)
    else:
        assert rg.result(0) is None
    rg.close()
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 4])
def test_two_ranks_shard_and_gather(tmp_path, n_pairs, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
    want = []
    for p in range(n_pairs):
        f1, f2 = make_pair(32, 48, 700 + p, shift=(p % 3 - 1, 1), patches=0, noise=0)
        flow, _ = oracle.estimate(f1, f2, [8], [4], 2)
        want.append(flow[::2, ::2].astype(np.int16))
    want = np.stack(want)
    for r in range(world):
        got = np.load(tmp_path / f"full_{r}.npy")
        assert got.shape == want.shape
        assert np.array_equal(got, want), r
