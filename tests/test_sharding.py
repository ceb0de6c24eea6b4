"""Multi-GPU path on CPU: contiguous shards, no data-path collective; gloo world_size 2 for the plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asr_ttl_mtl_b200.sharding import shard_range


def test_shards_tile_the_batch_in_rank_order():
    for n in (0, 1, 7, 256, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_config5_chunks_cover_every_clip_once():
    """bench.py --gpus N: rank r walks its shard of the 65,536 clips in 256-clip chunks with globally unique chunk indices."""
    import bench

    for world in (2, 4, 8):
        seen = []
        for rank in range(world):
            chunks = bench.shard_chunks(bench.TOTAL_CLIPS_CONFIG5, rank, world, 256)
            assert len(chunks) == bench.TOTAL_CLIPS_CONFIG5 // world // 256
            assert chunks[0][1] == shard_range(bench.TOTAL_CLIPS_CONFIG5, rank, world)[0]
            seen += [(g, first, n) for g, first, n in chunks]
        assert [g for g, _f, _n in seen] == list(range(256))
        assert sum(n for _g, _f, n in seen) == bench.TOTAL_CLIPS_CONFIG5
    odd = [c for r in range(3) for c in bench.shard_chunks(1000, r, 3, 256)]
    assert sum(n for _g, _f, n in odd) == 1000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    import sys

    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    import bench
    from oracle import logmel_oracle, signals

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = shard_range(n_clips, rank, world)
        # each rank computes its own utterances only; nothing is exchanged on the data path
        sums = [float(logmel_oracle.logmel_f32_port(signals.make_signal("gauss", 3200, 1000 + i), 80).double().sum())
                for i in range(begin, end)]
        elapsed_ms = 10.0 * (rank + 1)
        worst = bench.max_over_ranks(elapsed_ms, device="cpu")
        total = bench.sum_over_ranks(float(end - begin), device="cpu")
        gathered = [None] * world
        dist.all_gather_object(gathered, (begin, end, sums))
        if rank == 0:
            q.put((worst, total, gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_run_covers_the_batch_once():
    from oracle import logmel_oracle, signals

    world, n_clips = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    worst, total, gathered = q.get()
    assert worst == 20.0 and total == n_clips  # max-over-ranks timing, whole-job unit count
    assert [g[:2] for g in gathered] == [(0, 3), (3, 5)]
    sums = [s for g in gathered for s in g[2]]
    want = [float(logmel_oracle.logmel_f32_port(signals.make_signal("gauss", 3200, 1000 + i), 80).double().sum())
            for i in range(n_clips)]
    assert np.allclose(sums, want, rtol=0, atol=1e-6)  # same clip -> same result on whichever rank owns it
