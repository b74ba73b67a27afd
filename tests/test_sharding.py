"""Host logic of the multi-GPU partition: micro-batch plan, rank assignment, and the reporting
reduction under a real 2-process `gloo` group on CPU."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emojivoice_b200 import batch, sharding


def _lengths(n, seed=0):
    rng = random.Random(seed)
    return [2 * rng.randint(20, 150) + 1 for _ in range(n)]


def test_microbatches_cover_every_utterance_once_and_bound_padding():
    lens = _lengths(1024)
    mbs = sharding.microbatches(lens, 32)
    assert len(mbs) == 32 and all(len(m.items) == 32 for m in mbs)
    seen = sorted(i for m in mbs for i in m.items)
    assert seen == list(range(1024))
    for m in mbs:                                       # sorted by length: padding inside a batch stays small
        assert m.tx_max == max(lens[i] for i in m.items)
        assert m.tx_max - min(lens[i] for i in m.items) <= 12
    unsorted = sharding.microbatches(lens, 32, sort=False)
    assert [i for m in unsorted for i in m.items] == list(range(1024))      # the reference CLI's order (cli.py:281-286)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_assignment_is_deterministic_disjoint_and_balanced(world):
    lens = _lengths(1000, seed=3)
    plan_a = sharding.assign(sharding.microbatches(lens, 32), world)
    plan_b = sharding.assign(sharding.microbatches(lens, 32), world)
    assert [[m.index for m in r] for r in plan_a] == [[m.index for m in r] for r in plan_b]
    all_idx = sorted(m.index for r in plan_a for m in r)
    assert all_idx == list(range(32))                   # 1000 utterances -> 32 micro-batches (last one ragged: 8)
    loads = [sum(m.cost for m in r) for r in plan_a]
    assert max(loads) <= 1.25 * (sum(loads) / world)    # LPT keeps the slowest rank within 25 % of the mean
    for r in range(world):
        assert [m.index for m in sharding.shard(lens, 32, r, world)] == [m.index for m in plan_a[r]]


def test_edge_cases():
    assert sharding.microbatches([], 32) == []
    assert sharding.assign([], 4) == [[], [], [], []]
    one = sharding.microbatches([5], 32)
    assert len(one) == 1 and one[0].items == [0]
    with pytest.raises(ValueError):
        sharding.microbatches([1, 2], 0)
    with pytest.raises(ValueError):
        sharding.shard([1, 2], 2, rank=2, world_size=2)
    x, xl, spk = batch.collate([([1, 2, 3], 7), ([4], 12)], [1, 0])
    assert x.tolist() == [[4, 0, 0], [1, 2, 3]] and xl.tolist() == [1, 3] and spk.tolist() == [12, 7]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stats = sharding.ShardStats()
        mine = sharding.shard(lens, 16, rank, world, n_timesteps=10)
        for mb in mine:                                 # stand-in for synthesise+vocoder: 3 frames per token
            tx = [lens[i] for i in mb.items]
            stats.add([3 * t for t in tx], tx, 10, seconds=0.01 * len(tx) * (rank + 1))
        total = sharding.reduce_stats(stats)
        q.put((rank, sorted(i for mb in mine for i in mb.items), total))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_and_reporting_reduction():
    lens = _lengths(100, seed=5)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    items0, items1 = got[0][1], got[1][1]
    assert sorted(items0 + items1) == list(range(100)) and not set(items0) & set(items1)
    single = sharding.ShardStats()
    single.add([3 * t for t in lens], lens, 10, 0.0)
    for _, _, total in got:                             # every rank sees the same whole-job totals
        assert total["world_size"] == 2 and total["utterances"] == 100 and total["frames"] == single.frames
        assert abs(total["audio_seconds"] - single.audio_seconds) < 1e-6
        assert abs(total["flops"] - single.flops) < 1e-3 * single.flops
        assert abs(total["seconds"] - max(0.01 * len(items0), 0.02 * len(items1))) < 1e-9      # slowest rank bounds the job
