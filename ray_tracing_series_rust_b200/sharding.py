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


def path_range(n_pixels: int, spp: int, rank: int, world: int):
    """Even split of ANY sample count: rank r renders the path indices [T r / N, T (r + 1) / N) of the sample-major enumeration
    (index = sample * pixels + pixel, T = pixels * spp) through rt_render_device_paths: whole samples plus a partial first / last one
    (500 spp on 8 GPUs = 62.5 each instead of 62 or 63)."""
    if world < 1 or not (0 <= rank < world) or spp < 1 or n_pixels < 1:
        raise ValueError("bad shard request")
    total = n_pixels * spp
    return (total * rank) // world, (total * (rank + 1)) // world


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


class PeerGroup:
    """The accumulator exchange of the one-process-per-GPU layout done by librtb200 itself over NVLink peer memory (rt_peer_*,
    include/rtb200.h): every rank renders into a library-owned accumulator exported through CUDA IPC, rank 0 maps them all and one
    kernel there waits for each peer's published flag, sums the shards in place and resolves the Screen.  torch.distributed is used
    once, to hand the 64-byte IPC handles round (plumbing); the data path has no collective.

        pg = PeerGroup(api, rank, world, W * H * 3)           # collective: all ranks
        per step:  pg.begin(stream); rt_render_device(..., pg.accum, stream, NO_WAIT); pg.publish(stream)
                   rank 0: pg.gather_resolve(screen_ptr, W, H, spp, rows, stream)
    """

    def __init__(self, api, rank: int, world: int, n_elems: int):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.api, self.rank, self.world = api, rank, world
        handle = (C.c_uint8 * 64)()
        self.h = api.peer_create(rank, world, int(n_elems), handle)
        if not self.h:
            raise RuntimeError("rt_peer_create failed: " + api.err())
        mine = torch.tensor(list(handle), dtype=torch.uint8)
        if world > 1:
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            parts = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(parts, mine.to(dev))
            allh = torch.cat([p.cpu() for p in parts])
        else:
            allh = mine
        buf = (C.c_uint8 * (64 * world))(*allh.tolist())
        api.check(api.peer_connect(self.h, buf))
        self.accum = api.peer_accum(self.h)

    def begin(self, stream_ptr):
        self.api.check(self.api.peer_begin(self.h, stream_ptr))

    def publish(self, stream_ptr):
        self.api.check(self.api.peer_publish(self.h, stream_ptr))

    def gather_resolve(self, screen_ptr, W, H, spp, rows, stream_ptr):
        self.api.check(self.api.peer_gather_resolve(self.h, screen_ptr, W, H, spp, rows, stream_ptr))

    def timed_out(self) -> bool:
        return self.api.peer_timed_out(self.h) != 0

    def close(self):
        if self.h:
            self.api.peer_destroy(self.h)
            self.h = None


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
