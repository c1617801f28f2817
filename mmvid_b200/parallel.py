"""Multi-GPU sampling: replicas + batch split + ONE all-gather of decoded frames.

The path shards by independent units (each sample of the sampling batch is an independent sequence; the
reference itself loops over samples serially, dalle_bert.py:618), so there is no data-path collective besides
the final gather.  One process per GPU (torchrun); backend nccl on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous split of n samples over `world` ranks; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t, rank=None, world=None):
    if t is None:
        return None
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def all_gather_variable(x, sizes=None):
    """All-gather tensors whose leading dim may differ by rank (ragged batch split).  Pads to the max, gathers
    with a single collective, then trims.  Returns the concatenation in rank order on every rank."""
    world = dist.get_world_size()
    if sizes is None:
        n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.long)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n)
        sizes = [int(v) for v in all_n]
    m = max(sizes)
    if m == 0:
        return x
    pad = torch.zeros((m,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    pad[: x.shape[0]] = x
    out = torch.empty((world * m,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(out, pad.contiguous())
    chunks = [out[r * m: r * m + sizes[r]] for r in range(world)]
    return torch.cat(chunks, 0)


def _generate_per_sample(model, text, visual, seeds, gen_kwargs):
    """One sample at a time, each under its own seed: the result of sample i depends on (prompt i, seed i) only, so any
    partition of the batch over ranks reproduces the single-process output exactly."""
    imgs, seqs = [], []
    for i in range(text.shape[0]):
        torch.manual_seed(int(seeds[i]))
        im, _, sq = model.generate_images(text[i:i + 1], visual=None if visual is None else visual[i:i + 1], **gen_kwargs)
        imgs.append(im)
        seqs.append(sq)
    return torch.cat(imgs, 0), [], (torch.cat(seqs, 0) if seqs[0] is not None else None)


@torch.no_grad()
def generate_images_sharded(model, text, visual=None, sample_seeds=None, **gen_kwargs):
    """Global batch in, global batch out.  Rank r generates samples [lo_r, hi_r) with its replica and the decoded
    frames (and token ids) are all-gathered once at the end.  RNG: in throughput mode callers seed `seed + rank`
    (train.py:87 does the same) and the whole local shard is sampled at once; `sample_seeds` (one int per GLOBAL sample)
    selects partition-independent sampling instead: sample i is generated alone under seed sample_seeds[i], so the
    gathered ids / frames equal those of a single process given the same seeds."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        if sample_seeds is not None:
            return _generate_per_sample(model, text, visual, sample_seeds, gen_kwargs)
        return model.generate_images(text, visual=visual, **gen_kwargs)
    rank = dist.get_rank()
    n = text.shape[0]
    sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
    t_loc, v_loc = shard_batch(text, rank, world), shard_batch(visual, rank, world)
    if t_loc.shape[0] > 0 and sample_seeds is not None:
        lo, hi = shard_bounds(n, rank, world)
        images, extra, seq = _generate_per_sample(model, t_loc, v_loc, list(sample_seeds)[lo:hi], gen_kwargs)
    elif t_loc.shape[0] > 0:
        images, extra, seq = model.generate_images(t_loc, visual=v_loc, **gen_kwargs)
    else:
        raise RuntimeError("global batch smaller than world size")
    images = all_gather_variable(images.contiguous(), sizes)
    if seq is not None:
        per = seq.shape[0] // max(t_loc.shape[0], 1)
        seq = all_gather_variable(seq.contiguous(), [s * per for s in sizes])
    return images, extra, seq


@torch.no_grad()
def all_reduce_gradients(params, bucket_bytes=64 << 20):
    """Data-parallel training (train.py:32 wraps the model in DDP): average `.grad` over ranks with one all-reduce per
    flat bucket of up to `bucket_bytes` (few large NVLink/NVSwitch collectives instead of one per tensor), in parameter
    order so every rank builds identical buckets.  Parameters without a gradient on this rank contribute zeros."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    params = [p for p in params if p.requires_grad]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
