"""Sharding of independent frame pairs over the GPUs of one box, and the gather of results.

Frame pairs are independent units (one MF object per pair in the reference, main_class.cpp:45), so the path shards
by pair with no data-path collective; the only communication is the final gather of motion fields (NCCL over
NVLink on GPUs; the same code runs on gloo for the CPU tests).  Fields are gathered in the compact form (the
2x2-granular int16 field, 1/8 of the dense CV_32FC2 bytes) unless the caller passes dense tensors.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size):
    """Contiguous, balanced shards: returns [(start, stop)] per rank; earlier ranks take the remainder."""
    base, rem = divmod(int(n_items), int(world_size))
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def my_shard(n_items, rank=None, world_size=None):
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    return shard_bounds(n_items, world_size)[rank]


def gather_fields(local, n_total, dst=None, group=None):
    """Gathers per-pair result tensors.  local: [n_local, ...] on this rank's device, the pairs of my_shard(n_total).
    dst=None: every rank receives the full [n_total, ...] tensor (all_gather); dst=r: only rank r does (others get None).
    Shards may be ragged: they are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = shard_bounds(n_total, world)
    assert local.shape[0] == bounds[rank][1] - bounds[rank][0], "local tensor does not match this rank's shard"
    biggest = max(b - a for a, b in bounds)
    padded = local
    if local.shape[0] != biggest:
        padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    padded = padded.contiguous()
    # the collective moves raw bytes (int16 is not a collective dtype on every backend)
    elem_shape, elem_dtype = tuple(local.shape[1:]), local.dtype
    padded = padded.view(torch.uint8).reshape(biggest, -1)
    if dst is None:
        full = torch.empty((world * biggest, padded.shape[1]), dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(full, padded, group=group)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank != dst:
            return None
        full = torch.cat(parts, dim=0)
    full = full.view(elem_dtype).reshape((world * biggest,) + elem_shape)
    if all(b - a == biggest for a, b in bounds):
        return full
    return torch.cat([full[r * biggest:r * biggest + (b - a)] for r, (a, b) in enumerate(bounds)], dim=0)


class ResultGather:
    """Per-step gather of every rank's compact fields on rank 0 without putting communication kernels on the SMs.

    The search kernel is persistent (one CTA per SM, ~200 KB of shared memory each); an NCCL collective that overlaps it
    takes SMs away and the step waits for the CTAs that could not start (round 1: 5 % at 8 GPUs).  Frame pairs need no
    exchange at all, so the only transfer -- results to rank 0 -- goes through the copy engines: rank 0 allocates the
    gather buffer, shares it with the other ranks of the box as a CUDA IPC handle, and every rank copies its slice over
    NVLink with a device-to-device memcpy (no kernel).  NCCL (or gloo) carries the handle, the barriers and the timing
    reductions.  If the IPC mapping is not available, the fallback is `dist.gather` to rank 0 (send/recv kernels on a
    couple of channels); `how` says which one runs.
    """

    def __init__(self, like, n_total, n_buffers=2, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bounds = shard_bounds(n_total, self.world)
        self.n_local = self.bounds[self.rank][1] - self.bounds[self.rank][0]
        shape = (n_total,) + tuple(like.shape[1:])
        self.how = "nccl gather to rank 0"
        self.bufs = None       # rank 0: the gather buffers; other ranks: their IPC mappings (or None)
        self._parts = None
        self._like = like
        if like.is_cuda:
            try:
                from torch.multiprocessing.reductions import reduce_tensor
                payload = [None]
                if self.rank == 0:
                    self.bufs = [torch.empty(shape, dtype=like.dtype, device=like.device) for _ in range(n_buffers)]
                    payload = [[reduce_tensor(b) for b in self.bufs]]
                dist.broadcast_object_list(payload, src=0, group=group)
                ok = 1
                if self.rank != 0:
                    try:
                        self.bufs = [fn(*args) for fn, args in payload[0]]
                    except Exception:
                        ok = 0
                flag = torch.tensor([ok], dtype=torch.int32, device=like.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                if int(flag.item()) == 1:
                    self.how = "copy-engine peer copies over NVLink into rank 0's buffer (CUDA IPC), no communication kernel"
                elif self.rank != 0:
                    self.bufs = None
            except Exception:
                if self.rank != 0:
                    self.bufs = None
        if self.rank == 0 and self.bufs is None:
            self.bufs = [torch.empty(shape, dtype=like.dtype, device=like.device) for _ in range(n_buffers)]
        self.ipc = self.how.startswith("copy-engine")

    def push(self, local, which=0):
        """Enqueue this rank's slice of gather buffer `which` on the current stream."""
        a, b = self.bounds[self.rank]
        assert local.shape[0] == b - a
        if self.ipc or (self.rank == 0 and self.world == 1):
            self.bufs[which][a:b].copy_(local, non_blocking=True)
            return
        flat = local.contiguous().view(torch.uint8).reshape(local.shape[0], -1)
        if self.rank == 0:
            parts = [torch.empty_like(flat) for _ in range(self.world)]
            dist.gather(flat, parts, dst=0, group=self.group)
            for r, (ra, rb) in enumerate(self.bounds):
                self.bufs[which][ra:rb].view(torch.uint8).reshape(rb - ra, -1).copy_(parts[r][:rb - ra])
        else:
            dist.gather(flat, None, dst=0, group=self.group)

    def result(self, which=0):
        """Rank 0, after a barrier that follows every rank's push (and a device synchronisation): the gathered tensor."""
        return self.bufs[which] if self.rank == 0 else None

    def close(self):
        self.bufs = None
