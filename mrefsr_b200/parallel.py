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

* **ragged reference groups** (BASELINE config 3: LMR-shaped groups with 2-6 references per image) -- still batch
  sharding, but images cost differently and cannot all share one launch: `shard_ragged` balances images over the
  ranks by a cost model (no communication: every rank computes the same assignment), `batches_by_shape` /
  `run_ragged` bucket a rank's images by (reference count, size) so that each bucket is one batched forward.

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


def image_cost(n_refs, pixels=1.0):
    """Relative cost of one image of the x4 network: a part that does not depend on the references (content
    extractor, residual trunk, upsampling: ~1/3 of a 5-reference image in profiles/r01_full_model.md) plus a part per
    reference (VGG / extractor features, matcher, offset convolutions, DCN, fusion inputs); both scale with the
    pixel count, the matcher with its square (N_lr x N_ref correlations)."""
    return pixels * (1.0 + 0.4 * n_refs) + 0.05 * n_refs * pixels * pixels


def shard_ragged(ref_counts, world, rank=None, pixels=None):
    """Balanced assignment of images with different reference counts (and optionally sizes) to ranks.

    Longest-processing-time-first on `image_cost`: images sorted by cost (ties by index), each given to the least
    loaded rank (ties to the lowest rank).  Deterministic, so every rank derives the same assignment locally.
    Returns the sorted image indices of `rank`, or the list for all ranks when rank is None."""
    n = len(ref_counts)
    px = [1.0] * n if pixels is None else [float(p) for p in pixels]
    if pixels is not None and len(px) != n:
        raise ValueError('shard_ragged: %d pixel counts for %d images' % (len(px), n))
    if world < 1:
        raise ValueError('shard_ragged: world size must be positive')
    base = min(px) if px else 1.0
    cost = [image_cost(int(r), p / base) for r, p in zip(ref_counts, px)]
    order = sorted(range(n), key=lambda i: (-cost[i], i))
    load = [0.0] * world
    owned = [[] for _ in range(world)]
    for i in order:
        k = min(range(world), key=lambda q: (load[q], q))
        owned[k].append(i)
        load[k] += cost[i]
    owned = [sorted(o) for o in owned]
    return owned if rank is None else owned[rank]


def batches_by_shape(keys, max_batch):
    """Bucket item indices by key (e.g. (n_refs, H, W)), order-preserving, at most max_batch per bucket.
    Returns a list of (key, [indices])."""
    if max_batch < 1:
        raise ValueError('batches_by_shape: max_batch must be positive')
    open_buckets, out = {}, []
    for i, k in enumerate(keys):
        b = open_buckets.get(k)
        if b is None or len(b) == max_batch:
            b = []
            open_buckets[k] = b
            out.append((k, b))
        b.append(i)
    return out


def run_ragged(forward, samples, max_batch=16):
    """Run `forward(lq [B,3,h,w], up [B,3,H,W], refs [B,R,3,H,W]) -> [B,3,H,W]` (e.g. MRefSRPipeline) over samples
    with different reference counts / sizes: sample i = (lq [3,h,w], up [3,H,W], refs [R_i,3,H,W]).  Samples that
    share (R, H, W) go through one batched call; results come back in input order."""
    keys = []
    for lq, up, refs in samples:
        if refs.dim() != 4 or tuple(refs.shape[-2:]) != tuple(up.shape[-2:]):
            raise ValueError('run_ragged: refs must be [R,3,H,W] with the size of the upsampled input')
        keys.append((int(refs.shape[0]),) + tuple(up.shape[-2:]) + tuple(lq.shape[-2:]))
    out = [None] * len(samples)
    for _, idxs in batches_by_shape(keys, max_batch):
        lq = torch.stack([samples[i][0] for i in idxs])
        up = torch.stack([samples[i][1] for i in idxs])
        refs = torch.stack([samples[i][2] for i in idxs])
        sr = forward(lq, up, refs)
        if sr.shape[0] != len(idxs):
            raise RuntimeError('run_ragged: forward returned %d images for a batch of %d' % (sr.shape[0], len(idxs)))
        for j, i in enumerate(idxs):
            out[i] = sr[j]
    return out


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
        self.ptrs_by_rank = ptrs
        if len(self.ptrs) > 8:
            raise RuntimeError('PeerGatherBuffer: at most 8 ranks per group')

    def begin(self):
        self.hdl.barrier(channel=0)

    def finish(self):
        self.hdl.barrier(channel=1)
        return self.buf


class PeerSlabBuffer(PeerGatherBuffer):
    """Pixel-slab variant of the exchange: every GPU owns H / world consecutive output rows and its buffer
    [n, R, C, H / world, W] receives, from every rank's DCN epilogue, those rows of that rank's references
    (dcn.dynagg_dcn_forward_into(..., out_ptrs=buf.ptrs_by_rank, slab_rows=buf.slab_rows)).  Each GPU then fuses its
    slab over all R references and `all_gather_slabs` reassembles the fused result: 1/world of the fusion work and
    ~1/world of the exchanged bytes of the all-gather of aligned features.

        buf = PeerSlabBuffer(n, R, C, H, W, device)
        buf.begin(); dynagg_dcn_forward_into(..., buf.ptrs_by_rank, r_local, R, lo, slab_rows=buf.slab_rows)
        slab = buf.finish()              # [n, R, C, H / world, W]: all references, this rank's rows
    """

    def __init__(self, n, n_refs, channels, height, width, device, group=None):
        if not dist.is_initialized():
            raise RuntimeError('PeerSlabBuffer needs an initialised process group')
        world = dist.get_world_size(group)
        if height % world:
            raise RuntimeError('PeerSlabBuffer: %d rows do not split evenly over %d ranks' % (height, world))
        self.slab_rows = height // world
        super().__init__((n, n_refs, channels, self.slab_rows, width), device, group)


def slab_of(t, rank, world, dim=2):
    """This rank's consecutive rows of a full-height tensor (dim = the height axis)."""
    rows = t.shape[dim] // world
    return t.narrow(dim, rank * rows, rows).contiguous()


def all_gather_slabs(slab, group=None):
    """[n, C, H / world, W] per rank -> [n, C, H, W] on every rank (rows in rank order)."""
    if not dist.is_available() or not dist.is_initialized():
        return slab
    world = dist.get_world_size(group)
    n, c, hs, w = slab.shape
    recv = slab.new_empty((world, n, c, hs, w))
    if slab.is_cuda:
        dist.all_gather_into_tensor(recv.view(-1), slab.contiguous().view(-1), group=group)
    else:
        parts = [torch.empty_like(slab) for _ in range(world)]
        dist.all_gather(parts, slab.contiguous(), group=group)
        recv = torch.stack(parts, 0)
    return recv.permute(1, 2, 0, 3, 4).reshape(n, c, world * hs, w)
