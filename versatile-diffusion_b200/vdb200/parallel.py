"""Multi-GPU plumbing of the sampling path: one process per GPU, the batch is sharded, weights replicated.

Every sample is independent end to end (GroupNorm/LayerNorm are per-sample, CFG pairs stay on one GPU), so the
only exchange is the context-embedding broadcast at the start of sampling and an optional
gather of the decoded images.  Works on any torch.distributed backend (NCCL over NVLink on the B200 box, gloo
in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(global_batch, rank=None, world_size=None):
    """Rows [lo, hi) of the global batch owned by `rank` (contiguous, near-equal shards)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def seeded_latents(rows, shape, seed, dtype=torch.float32):
    """x_T rows generated per GLOBAL row index, so an N-rank run draws exactly the rows of the 1-rank run."""
    lo, hi = rows
    out = torch.empty((hi - lo,) + tuple(shape), dtype=dtype)
    for i in range(lo, hi):
        g = torch.Generator().manual_seed((seed * 1000003 + i) & 0x7FFFFFFFFFFFFFFF)
        out[i - lo] = torch.randn(shape, generator=g, dtype=dtype)
    return out


def broadcast_context(tensors, src=0):
    """Rank `src` holds the encoded contexts (cond / uncond embeddings); every other rank receives them in place."""
    _, w = world()
    if w > 1:
        for t in tensors:
            dist.broadcast(t, src)
    return tensors


def gather_images(images, dst=0):
    """Concatenate the per-rank image shards on `dst` (None elsewhere). Shards may differ in length by one row."""
    r, w = world()
    if w == 1:
        return images
    sizes = [torch.zeros(1, dtype=torch.long, device=images.device) for _ in range(w)]
    dist.all_gather(sizes, torch.tensor([images.shape[0]], dtype=torch.long, device=images.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    pad[:images.shape[0]] = images
    bufs = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(bufs, pad)
    if r != dst:
        return None
    return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)], dim=0)
