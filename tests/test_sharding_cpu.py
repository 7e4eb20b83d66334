"""Host logic of the multi-GPU path on CPU: contiguous batch sharding and the (test-only) feature gather,
exercised with world_size 2 and 3 over the gloo backend."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from panoswintransformerobjectdetection_b200.runtime import gather_features, shard_batch, shard_bounds


def test_shard_bounds_cover_the_batch_exactly():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    batch = torch.randn(n_items, 3, 8, 16)                     # same global batch on every rank
    mine = shard_batch(batch, rank, world)
    # stand-in for the per-image backbone: any per-image function commutes with sharding
    feats = [mine.mean(1, keepdim=True) * 2.0, mine[:, :, ::2, ::2].contiguous()]
    full = gather_features(feats, world)
    want = [batch.mean(1, keepdim=True) * 2.0, batch[:, :, ::2, ::2]]
    ok = all(torch.equal(a, b) for a, b in zip(full, want))
    q.put((rank, ok, mine.shape[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 8), (2, 5), (3, 7)])
def test_sharded_forward_reassembles_over_gloo(world, n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + n_items
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == n_items
