"""Multi-GPU layouts of the alignment hot path (one process per GPU, torch.distributed; NCCL on GPUs).

Two modes, as SURVEY.md section 8e derives from the reference's data flow:

* **batch sharding** -- images are independent end to end (the matcher loops per batch item,
  basicsr/archs/corres_generation_arch.py:53; DCN loops per sample, deform_conv_cuda.cpp:539), so ranks take
  contiguous batch slices and run the whole path with no data-path collective.  `shard_range` / `shard_batch`.
* **reference sharding** -- for many / large references (BASELINE config 4) the refs of an image are split over
  the ranks: extractor -> matcher -> pre-offsets -> DCN are independent per reference *within a scale*; the
  references meet only in MRAPAFusion's softmax over t (ref_mrapa_restoration_arch.py:331-333), once per scale.
  Each rank aligns its own references and `all_gather_refs` exchanges the aligned features (one NCCL all-gather
  per scale); every rank then runs the fusion on the full set.

Only the exchange lives here; the kernels are the single-GPU ones.  Uneven splits (R not divisible by the world
size) are padded to the largest shard for the collective and trimmed afterwards.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra element."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def shard_batch(tensors, rank, world, dim=0):
    """Slice every tensor of a dict / list / single tensor along `dim` for this rank (batch sharding)."""
    def cut(t):
        lo, hi = shard_range(t.shape[dim], rank, world)
        return t.narrow(dim, lo, hi - lo)
    if isinstance(tensors, dict):
        return {k: cut(v) for k, v in tensors.items()}
    if isinstance(tensors, (list, tuple)):
        return type(tensors)(cut(v) for v in tensors)
    return cut(tensors)


def all_gather_refs(local, n_refs_total, group=None):
    """Exchange per-reference aligned features.

    local: [n, r_local, C, h, w] -- this rank's references (r_local = shard of n_refs_total, contiguous split).
    returns [n, n_refs_total, C, h, w], references in global order, identical on every rank.
    One collective; shards are padded to the largest r_local so that all_gather sees equal sizes.
    """
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[1] != n_refs_total:
            raise RuntimeError('all_gather_refs: no process group but local shard is not the full reference set')
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(n_refs_total, world)
    rank = dist.get_rank(group)
    if local.shape[1] != counts[rank]:
        raise RuntimeError('all_gather_refs: rank %d holds %d references, expected %d'
                           % (rank, local.shape[1], counts[rank]))
    rmax = max(counts)
    n = local.shape[0]
    tail = tuple(local.shape[2:])
    send = local.contiguous()
    if counts[rank] < rmax:
        pad = local.new_zeros((n, rmax - counts[rank]) + tail)
        send = torch.cat((send, pad), dim=1).contiguous()
    # gather as [world, n, rmax, ...]
    recv = local.new_empty((world, n, rmax) + tail)
    if send.is_cuda:
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    else:  # gloo (CPU tests) has no all_gather_into_tensor for arbitrary layouts on every version
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        recv = torch.stack(parts, 0)
    pieces = [recv[r, :, :counts[r]] for r in range(world) if counts[r] > 0]
    return torch.cat(pieces, dim=1)


def align_reference_sharded(align_fn, refs_global, n_refs_total, group=None):
    """Reference-sharded alignment of one scale.

    align_fn(ref_index) -> aligned feature [n, C, h, w] for one global reference index (runs the single-GPU
    matcher / DynAgg path for that reference); refs_global is the list of global reference indices owned by this
    rank (its `shard_range`).  Returns the full [n, n_refs_total, C, h, w] on every rank.
    """
    feats = [align_fn(r) for r in refs_global]
    if feats:
        local = torch.stack(feats, dim=1)
    else:
        raise RuntimeError('align_reference_sharded: a rank without references needs a template shape; '
                           'use world_size <= n_refs')
    return all_gather_refs(local, n_refs_total, group)


class PeerGatherBuffer:
    """The gathered tensor [n, R, C, H, W] of the reference-sharded mode as symmetric memory: one copy per GPU, every
    copy mapped into every process over NVLink (torch.distributed._symmetric_memory).  The DCN epilogue of each rank
    stores its references' aligned features into ALL copies (dcn.dynagg_dcn_forward_into), so the exchange overlaps
    the kernel tile by tile and there is no transpose copy, no ncclAllGather and no cat afterwards.

        buf = PeerGatherBuffer((n, R, C, H, W), device)
        buf.begin()                      # peers have finished reading the previous contents
        dynagg_dcn_forward_into(..., out_ptrs=buf.ptrs, dst_group=r_local, dst_stride=R, dst_offset=lo)
        full = buf.finish()              # all ranks' stores have landed -> [n, R, C, H, W] on this GPU
    """

    def __init__(self, shape, device, group=None):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError('PeerGatherBuffer needs an initialised process group')
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(*shape, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        rank = dist.get_rank(self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptrs = [ptrs[rank]] + [p for i, p in enumerate(ptrs) if i != rank]     # local copy first
        if len(self.ptrs) > 8:
            raise RuntimeError('PeerGatherBuffer: at most 8 ranks per group')

    def begin(self):
        self.hdl.barrier(channel=0)

    def finish(self):
        self.hdl.barrier(channel=1)
        return self.buf
