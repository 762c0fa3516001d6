"""N>1 host path on CPU: two gloo ranks render disjoint sample ranges (with the oracle standing in for
the GPU renderer — same C-ABI, same rt_render_config sample_begin/sample_end contract), the int64
accumulators are summed with one reduce, and rank 0's result equals the single-process render."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scaling, out_path, tiles=False, how="reduce"):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from ray_tracing_series_rust_b200 import capi, sharding
    s = oracle.new_scene()
    s.world_build(5, 0xB002)
    s.commit()
    spp = 6
    if tiles:  # tile sharding: every rank renders all samples of its own 4-row bands
        total, b, e = spp, 0, spp
        cfg = capi.make_config(40, 1.0, total, 50, seed=3, threads=2, flags=sharding.tile_flags(rank, world))
    else:
        total, b, e = sharding.sample_range(spp, rank, world, scaling)
        cfg = capi.make_config(40, 1.0, total, 50, seed=3, sample_begin=b, sample_end=e, threads=2)
    _, acc, st = s.render(cfg, want_accum=True)
    t = torch.from_numpy(acc)
    paths = torch.tensor([st["paths"]], dtype=torch.int64)
    sharding.reduce_accumulators(t, dst=0, how=how)
    dist.reduce(paths, dst=0, op=dist.ReduceOp.SUM)
    if how == "allreduce" and rank == 1:  # every rank ends with the sum
        np.save(out_path + ".rank1.npy", t.numpy())
    if rank == 0:
        np.save(out_path, t.numpy())
        np.save(out_path + ".paths.npy", paths.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scaling", ["strong", "weak"])
def test_two_rank_sample_sharding_matches_single_process(orc, tmp_path, scaling):
    from ray_tracing_series_rust_b200 import capi
    world = 2
    out = str(tmp_path / f"acc_{scaling}.npy")
    mp.spawn(_worker, args=(world, _free_port(), scaling, out), nprocs=world, join=True)
    got = np.load(out)
    paths = int(np.load(out + ".paths.npy")[0])
    total = 6 * world if scaling == "weak" else 6
    s = orc.new_scene()
    s.world_build(5, 0xB002)
    s.commit()
    _, ref, st = s.render(capi.make_config(40, 1.0, total, 50, seed=3, threads=2), want_accum=True)
    assert paths == st["paths"] == 40 * 40 * total
    # the oracle rounds each shard's f64 sum to fixed point once, so shards differ from the whole by at most one unit per rank
    assert np.abs(got - ref).max() <= world


def test_two_rank_tile_sharding_matches_single_process(orc, tmp_path):
    from ray_tracing_series_rust_b200 import capi, sharding
    world = 2
    out = str(tmp_path / "acc_tiles.npy")
    mp.spawn(_worker, args=(world, _free_port(), "strong", out, True), nprocs=world, join=True)
    got = np.load(out)
    s = orc.new_scene()
    s.world_build(5, 0xB002)
    s.commit()
    _, ref, st = s.render(capi.make_config(40, 1.0, 6, 50, seed=3, threads=2), want_accum=True)
    assert int(np.load(out + ".paths.npy")[0]) == st["paths"] == 40 * 40 * 6
    assert np.array_equal(got, ref)  # disjoint pixels: the sum is exact
    # one shard alone: its bands equal the full render, the other rows are untouched zeros
    _, part, stp = s.render(capi.make_config(40, 1.0, 6, 50, seed=3, threads=2, flags=sharding.tile_flags(1, 3)), want_accum=True)
    mine = sharding.tile_rows(40, 1, 3)
    other = [j for j in range(40) if j not in mine]
    assert np.array_equal(part[mine], ref[mine]) and not part[other].any() and stp["paths"] == len(mine) * 40 * 6
    assert sorted(sum((sharding.tile_rows(41, r, 3) for r in range(3)), [])) == list(range(41))
    with pytest.raises(capi.RtError):
        s.render(capi.make_config(40, 1.0, 6, 50, flags=(5 << 16) | (3 << 24)))
    with pytest.raises(ValueError):
        sharding.tile_flags(3, 3)


def test_two_rank_allreduce_leaves_the_sum_on_every_rank(orc, tmp_path):
    """bench.py --reduce allreduce (NVLS-capable collective on NVSwitch boxes): same sum, on both ranks"""
    world = 2
    out = str(tmp_path / "acc_all.npy")
    mp.spawn(_worker, args=(world, _free_port(), "strong", out, False, "allreduce"), nprocs=world, join=True)
    assert np.array_equal(np.load(out), np.load(out + ".rank1.npy")) and np.load(out).any()


def test_sample_range_partition():
    from ray_tracing_series_rust_b200 import sharding
    for spp in (1, 7, 500, 10000):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                tot, b, e = sharding.sample_range(spp, r, world, "strong")
                assert tot == spp and 0 <= b <= e <= spp
                cover += list(range(b, e))
            assert cover == list(range(spp))
            tot, b, e = sharding.sample_range(spp, world - 1, world, "weak")
            assert (tot, e - b, e) == (spp * world, spp, spp * world)
    with pytest.raises(ValueError):
        sharding.sample_range(10, 2, 2)


def test_path_range_partition():
    """sharding.path_range: N contiguous, disjoint, near-equal ranges of the sample-major path enumeration, for any spp / N"""
    from ray_tracing_series_rust_b200 import sharding
    for npix, spp in ((426400, 500), (1000000, 10000), (7, 3), (1, 1)):
        for world in (1, 2, 3, 8):
            edges = [sharding.path_range(npix, spp, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == npix * spp
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.path_range(426400, 500, 3, 8) == (426400 * 500 * 3 // 8, 426400 * 500 * 4 // 8)  # 62.5 spp each
    with pytest.raises(ValueError):
        sharding.path_range(10, 5, 2, 2)


def test_frames_for_rank_partition():
    from ray_tracing_series_rust_b200 import sharding
    for n in (0, 1, 7, 240):
        for world in (1, 2, 4, 8):
            got = sorted(f for r in range(world) for f in sharding.frames_for_rank(n, r, world))
            assert got == list(range(n))
    assert sharding.frames_for_rank(240, 3, 8)[:3] == [3, 11, 19]
