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
