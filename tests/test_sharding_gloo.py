"""world_size = 2 over gloo on CPU: the N > 1 host logic (shard ranges, id broadcast) and the property the
multi-GPU path relies on — partial images of the ray shards sum to the full image (here with the oracle as the
per-rank renderer, since there is no GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    import lg_oracle as oracle
    from light_garden_b200 import abi, scenes
    from light_garden_b200.distributed import broadcast_bytes, shard_count, shard_indices
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. id broadcast helper
        secret = bytes(range(128)) if rank == 0 else b""
        got = broadcast_bytes(secret, 128, rank, 0)
        assert got == bytes(range(128))
        # 2. the shards partition [0, n) without gaps or overlap, for awkward n
        for n in (0, 1, 7, 1000, 12345, 2 ** 40 + 3):
            cnt = shard_count(n, rank, world)
            t = torch.tensor([cnt], dtype=torch.int64)
            parts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(parts, t)
            assert sum(int(p[0]) for p in parts) == n
            if n <= 12345:
                mine = torch.zeros(max(n, 1), dtype=torch.int64)
                mine[list(shard_indices(n, rank, world))] = 1
                dist.all_reduce(mine)
                assert n == 0 or bool((mine[:n] == 1).all())
        # 3. partial images of the shards sum to the whole frame
        spec = scenes.c1_default(total_rays=3000, width=240, height=135)
        for l in spec.lights:                                  # power-of-two colours: exact sums in any order
            l.color = (2.0 ** -7, 2.0 ** -8, 2.0 ** -7, 2.0 ** -6)
        osc = oracle.OracleScene.from_spec(spec)
        part = osc.trace_all(spec.lights, abi.LG_PRECISION_F32, rank=rank, world=world)
        assert part.primary_rays == sum(shard_count(l.num_rays, rank, world) for l in spec.lights)
        img = oracle.new_image(spec.width, spec.height, 1.0 if rank == 0 else 0.0)   # only the root owns the clear
        oracle.accumulate_segments(img, part.seg)
        t = torch.from_numpy(img)
        dist.reduce(t, 0, op=dist.ReduceOp.SUM)
        if rank == 0:
            full = osc.trace_all(spec.lights, abi.LG_PRECISION_F32)
            ref = oracle.new_image(spec.width, spec.height)
            oracle.accumulate_segments(ref, full.seg)
            # Fresnel-split colours are not powers of two: the two summation orders differ by fp32 rounding only
            assert np.array_equal(t.numpy()[..., 3] > 1, ref[..., 3] > 1)
            # two interleaved partial sums + one add vs one sequential fp32 sum per pixel (thousands of terms at the
            # hot pixels): bounded at 1e-4 of the pixel value
            assert (np.abs(t.numpy() - ref) <= 1e-4 * np.maximum(1.0, np.abs(ref))).all()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_count_partitions():
    from light_garden_b200.distributed import shard_count, shard_indices
    with pytest.raises(ValueError):
        shard_count(10, 2, 2)
    n = 33_333_333
    assert sum(shard_count(n, r, 8) for r in range(8)) == n
    assert sorted(i for r in range(3) for i in shard_indices(10, r, 3)) == list(range(10))
    assert shard_count(2, 5, 8) == 0
