"""N>1 host logic on CPU: world_size-2 gloo process group, contiguous batch shards, max-over-ranks
timing and a checksum-of-checksums that must not depend on how the batch was split."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    import torch.distributed as dist

    from chowdsp_fft_b200.sharding import max_over_ranks, shard_range, sum_over_ranks

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard_range(total, rank, world)
    # stand-in for the per-rank work: a checksum of this rank's slice of a deterministic batch
    rng = np.random.default_rng(42)
    batch = rng.uniform(-1, 1, (total, 64)).astype(np.float32)
    local = float(batch[start:start + count].astype(np.float64).sum())
    total_sum = sum_over_ranks(local)
    slowest = max_over_ranks(10.0 + rank)  # rank r "took" 10 + r ms
    dist.barrier()
    q.put((rank, start, count, total_sum, slowest))
    dist.destroy_process_group()


def test_shard_range_tiles_the_batch():
    from chowdsp_fft_b200.sharding import shard_range

    for total in (0, 1, 7, 934, 4096, 65536, 956416):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp

    world, total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    want = float(rng.uniform(-1, 1, (total, 64)).astype(np.float32).astype(np.float64).sum())
    assert results[0][1] == 0 and results[0][1] + results[0][2] == results[1][1]
    assert results[1][1] + results[1][2] == total
    for r in results:
        assert abs(r[3] - want) < 1e-6 * max(1.0, abs(want))  # checksum independent of the split
        assert r[4] == 11.0                                   # max over ranks
