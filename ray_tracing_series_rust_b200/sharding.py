"""Multi-GPU sharding of one render: every (pixel, sample) is independent
(src/world.rs:1208-1216 has no cross-iteration state but the per-pixel sum), so rank r of R renders the
sample range [begin, end) of every pixel and the int64 fixed-point accumulators are summed with ONE
reduce.  Philox streams are keyed by the global sample index, integer addition is associative: the
reduced image does not depend on R.

The reference's analogue is the row-band fan-out + mpsc gather of render_scene (world.rs:1198-1240).
"""
from __future__ import annotations


def sample_range(spp: int, rank: int, world: int, scaling: str = "strong"):
    """-> (spp_total, begin, end).  strong: the workload's spp is split across ranks (remainder spread
    over the first ranks); weak: every rank renders `spp` samples of an image with spp * world samples."""
    if world < 1 or not (0 <= rank < world) or spp < 1:
        raise ValueError("bad shard request")
    if scaling == "weak":
        return spp * world, spp * rank, spp * (rank + 1)
    if scaling != "strong":
        raise ValueError("scaling must be 'weak' or 'strong'")
    return spp, (spp * rank) // world, (spp * (rank + 1)) // world


def reduce_accumulators(accum, dst: int = 0, how: str = "reduce"):
    """Sum the per-rank int64 accumulators onto `dst` (NCCL on GPU tensors, gloo on CPU tensors).
    how = "reduce" (one ncclReduce to dst) or "allreduce" (every rank ends with the sum; on NVSwitch boxes NCCL can then reduce
    inside the switch, NVLS)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if how == "allreduce":
            dist.all_reduce(accum, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def frames_for_rank(n_frames: int, rank: int, world: int):
    """Animation sharding: frame f goes to rank f mod world; no collective, each frame is
    committed (shutter window) and rendered on one GPU.  [ref: render_scene_with_time, src/world.rs:1249]"""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request")
    return list(range(rank, n_frames, world))


TILE_ROWS = 4  # RT_TILE_ROWS (include/rtb200.h)


def tile_flags(rank: int, world: int) -> int:
    """rt_render_config.flags bits of RT_RENDER_TILE_SHARD(rank, world): rank renders the 4-row bands b with
    b % world == rank (SURVEY.md 8e "tile sharding", the GPU analogue of the row bands of world.rs:1198-1227).
    The shards are disjoint, so the image is their sum (same reduce as sample sharding) or a gather of the bands."""
    if world < 1 or world > 127 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    return 0 if world == 1 else ((rank & 0x7F) << 16) | ((world & 0x7F) << 24)


def tile_rows(height: int, rank: int, world: int):
    """Row indices (Screen rows, 0 = bottom) rank owns under tile sharding."""
    return [j for j in range(height) if (j // TILE_ROWS) % world == rank]
