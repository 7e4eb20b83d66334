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


def _train_worker(rank, world, port, q):
    """GraphedTrainStep (eager mode) over gloo: two ranks, each with half of a batch, must end up with the weights of
    one process trained on the whole batch (gradients averaged over equal shards = gradient of the mean loss)."""
    from panoswintransformerobjectdetection_b200.runtime import GraphedTrainStep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                              # DIFFERENT initial weights per rank: the broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    unused = torch.nn.Parameter(torch.ones(4))                 # never receives a gradient: must stay out of the optimizer
    net.register_parameter("unused", unused)
    torch.manual_seed(7)
    data = torch.randn(3, 8, 6)                                # three global batches of 8 rows
    ts = GraphedTrainStep(net, lambda o: o.square().mean(), shard_batch(data[0], rank, world),
                          lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9), graphs=False)
    for b in range(3):
        ts.step(shard_batch(data[b], rank, world))
    q.put((rank, [p.detach().tolist() for p in net.parameters()], len(ts.params)))   # lists: the worker exits before the parent reads
    dist.destroy_process_group()


def test_graphed_train_step_averages_gradients_over_gloo():
    from panoswintransformerobjectdetection_b200.runtime import GraphedTrainStep
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, 29688, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    # single-process reference: rank 0's initial weights (seed 100), whole batches
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net.register_parameter("unused", torch.nn.Parameter(torch.ones(4)))
    torch.manual_seed(7)
    data = torch.randn(3, 8, 6)
    ref = GraphedTrainStep(net, lambda o: o.square().mean(), data[0], lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9),
                           graphs=False)
    for b in range(3):
        ref.step(data[b])
    want = [p.detach() for p in net.parameters()]
    assert res[0][2] == res[1][2] == len(ref.params) == 4      # the unused parameter is not optimised
    for got in (res[0][1], res[1][1]):
        assert all(torch.allclose(torch.tensor(a), b, rtol=1e-5, atol=1e-6) for a, b in zip(got, want))
    assert torch.equal(dict(net.named_parameters())["unused"].detach(), torch.ones(4))
