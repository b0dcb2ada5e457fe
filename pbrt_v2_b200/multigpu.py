"""Image-tile sharding of one Render() across GPUs (one process per GPU).

pbrt-v2's render tasks share nothing (own RNG stream, own pixels), so the path
shards without any data-path collective: rank r renders the tasks whose block
index is congruent to r, and the per-rank framebuffers are summed once at the end
(NCCL reduce to rank 0; disjoint pixels add zeros, the rare edge splats add up).
Blocks of a few pixel rows are interleaved so that cheap (background) and
expensive (object) rows are spread evenly.
"""


def shard_render_desc(rd, rank, world, xres, rows_per_block=4):
    """Fill the shard fields of a RenderDesc for `rank` of `world` (returns rd)."""
    rd.shardCount = int(world)
    rd.shardIndex = int(rank)
    rd.shardBlock = max(1, int(rows_per_block) * int(xres))
    return rd


def reduce_film(film, world, dst=0):
    """Sum the per-rank films (a torch tensor, CUDA+NCCL or CPU+gloo) onto rank `dst`."""
    if world > 1:
        import torch.distributed as dist
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
