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
    gather buffers and exports them as CUDA IPC handles (libbbme: bbme_ipc_export), every other rank maps them with its own
    device current (bbme_ipc_open: peer access over NVLink, no context on rank 0's GPU) and pushes its slice with a
    device-to-device memcpy on its own stream (bbme_copy_async: no kernel).  NCCL (or gloo) carries the 64-byte handles, the
    barriers and the timing reductions.  If the mapping is not available (CPU tensors, no peer access), the fallback is
    `dist.gather` to rank 0; `how` says which one runs.
    """

    def __init__(self, like, n_total, n_buffers=2, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bounds = shard_bounds(n_total, self.world)
        self.shape = (n_total,) + tuple(like.shape[1:])
        self.dtype = like.dtype
        self.item_bytes = like[0].numel() * like.element_size() if like.shape[0] else 0
        self.how = "nccl gather to rank 0"
        self.ipc = False
        self.ptrs = []        # raw device pointers of the gather buffers (rank 0: owned, others: IPC mappings)
        self.bufs = None      # fallback path: torch tensors on rank 0
        self._lib = None
        self._dev = like.device.index if like.is_cuda else None
        if like.is_cuda:
            self._setup_ipc(like, n_total, n_buffers)
        if not self.ipc and self.rank == 0:
            self.bufs = [torch.empty(self.shape, dtype=like.dtype, device=like.device) for _ in range(n_buffers)]

    def _setup_ipc(self, like, n_total, n_buffers):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        nbytes = n_total * self.item_bytes
        ok = 1
        handles = torch.zeros((n_buffers, 64), dtype=torch.uint8)
        if self.rank == 0:
            for i in range(n_buffers):
                p = C.c_void_p()
                h = (C.c_ubyte * 64)()
                if lib.bbme_device_alloc(self._dev, nbytes, C.byref(p)) != 0 or lib.bbme_ipc_export(self._dev, p, h) != 0:
                    ok = 0
                    break
                self.ptrs.append(p.value)
                handles[i] = torch.frombuffer(bytearray(h), dtype=torch.uint8)
        hd = handles.to(like.device)
        dist.broadcast(hd, src=0, group=self.group)
        handles = hd.cpu()
        if self.rank != 0:
            for i in range(n_buffers):
                p = C.c_void_p()
                h = (C.c_ubyte * 64).from_buffer_copy(bytes(handles[i].tolist()))
                if lib.bbme_ipc_open(self._dev, h, C.byref(p)) != 0:
                    ok = 0
                    break
                self.ptrs.append(p.value)
        flag = torch.tensor([ok], dtype=torch.int32, device=like.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self._lib = lib
        if int(flag.item()) == 1:
            self.ipc = True
            self.how = "copy-engine peer copies over NVLink into rank 0's buffer (CUDA IPC mapping, cudaMemcpyAsync), no communication kernel"
        else:
            self._release()

    def push(self, local, which=0):
        """Enqueue this rank's slice of gather buffer `which` on the current stream."""
        a, b = self.bounds[self.rank]
        assert local.shape[0] == b - a
        if self.ipc:
            local = local.contiguous()
            stream = torch.cuda.current_stream(local.device).cuda_stream
            rc = self._lib.bbme_copy_async(self._dev, self.ptrs[which] + a * self.item_bytes, local.data_ptr(),
                                           (b - a) * self.item_bytes, stream)
            if rc != 0:
                raise RuntimeError("bbme_copy_async failed")
            return
        biggest = max(y - x for x, y in self.bounds)  # ragged shards are padded for the collective
        flat = torch.zeros((biggest, self.item_bytes), dtype=torch.uint8, device=local.device)
        flat[:b - a] = local.contiguous().view(torch.uint8).reshape(b - a, -1)
        if self.rank == 0:
            parts = [torch.empty_like(flat) for _ in range(self.world)]
            dist.gather(flat, parts, dst=0, group=self.group)
            for r, (ra, rb) in enumerate(self.bounds):
                self.bufs[which][ra:rb].view(torch.uint8).reshape(rb - ra, -1).copy_(parts[r][:rb - ra])
        else:
            dist.gather(flat, None, dst=0, group=self.group)

    def result(self, which=0):
        """Rank 0, after a barrier that follows every rank's push (and a device synchronisation): the gathered tensor."""
        if self.rank != 0:
            return None
        if not self.ipc:
            return self.bufs[which]
        out = torch.empty(self.shape, dtype=self.dtype, device=torch.device("cuda", self._dev))
        stream = torch.cuda.current_stream(out.device).cuda_stream
        if self._lib.bbme_copy_async(self._dev, out.data_ptr(), self.ptrs[which], out.numel() * out.element_size(), stream) != 0:
            raise RuntimeError("bbme_copy_async failed")
        torch.cuda.current_stream(out.device).synchronize()
        return out

    def _release(self):
        if self._lib is None:
            self.ptrs = []
            return
        for p in self.ptrs:
            if self.rank == 0:
                self._lib.bbme_device_free(self._dev, p)
            else:
                self._lib.bbme_ipc_close(self._dev, p)
        self.ptrs = []

    def close(self):
        """Every rank; the mappings first (call after a barrier), then the owner's buffers."""
        if self.ipc and self.rank != 0:
            self._release()
        if self.ipc:
            dist.barrier(group=self.group)
        if self.ipc and self.rank == 0:
            self._release()
        self.bufs = None
