"""world_size-2 gloo tests (CPU) of the batch-sharding / flat gradient all-reduce host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lamp_b200 import distributed as D


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(6, 5)
        self.dead = torch.nn.Linear(6, 5)      # never used in forward -> grad None (like the encoder self-attention)
        self.b = torch.nn.Linear(5, 3)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _worker(rank, world, port, uneven, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env('gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(1)
    n = 11 if uneven else 12
    x, y = torch.randn(n, 6), torch.randn(n, 3)
    model = Toy()
    xs, ys = D.shard_batch([x, y], rank, world)
    loss = torch.nn.functional.mse_loss(model(xs), ys)      # mean over the local shard
    loss.backward()
    weight = xs.shape[0] * world / n
    nel = D.allreduce_gradients(list(model.parameters()), local_weight=weight)
    ref = Toy()
    torch.nn.functional.mse_loss(ref(x), y).backward()      # single-process gradient on the whole batch
    ok = True
    for (name, p), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
        if pr.grad is None:
            ok &= p.grad is None
        else:
            ok &= torch.allclose(p.grad, pr.grad, rtol=1e-5, atol=1e-6)
    q.put((rank, bool(ok), nel))
    dist.destroy_process_group()


@pytest.mark.parametrize('uneven', [False, True])
def test_sharded_gradients_equal_single_process(uneven):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, uneven, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == sum(p.numel() for p in Toy().parameters())


def _reducer_worker(rank, world, port, uneven, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D.init_from_env('gloo')
    n = 11 if uneven else 12
    model, ref = Toy(), Toy()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1)
    red = D.GradientReducer(model.parameters(), bucket_mb=4e-5)   # >= 11 elements per bucket: two buckets
    ok = True
    for step in range(3):   # step 0 calibrates (blocking reduce), steps 1-2 reduce from the backward hooks
        g = torch.Generator().manual_seed(100 + step)
        x, y = torch.randn(n, 6, generator=g), torch.randn(n, 3, generator=g)
        xs, ys = D.shard_batch([x, y], rank, world)
        D.data_parallel_step(model, opt, lambda m, b: torch.nn.functional.mse_loss(m(b[0]), b[1]), (xs, ys),
                             reducer=red, local_n=xs.shape[0], global_n=n)
        opt_ref.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(ref(x), y).backward()
        for (name, p), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
            if pr.grad is None:
                ok &= p.grad is None                       # dead parameters stay grad-less, as in the reference
            else:
                ok &= torch.allclose(p.grad, pr.grad, rtol=1e-5, atol=1e-6)
                ok &= p.grad.data_ptr() >= red.flat.data_ptr()   # ... and live ones are views of the flat buffer
        opt_ref.step()
        for p, pr in zip(model.parameters(), ref.parameters()):
            ok &= torch.allclose(p, pr, rtol=1e-5, atol=1e-6)
    q.put((rank, bool(ok), dict(red.stats)))
    dist.destroy_process_group()


@pytest.mark.parametrize('uneven', [False, True])
def test_gradient_reducer_matches_single_process(uneven):
    """Flat-buffer bucketed reducer (hooks + async all-reduce) over 3 optimizer steps == single-process training."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, uneven, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    st = res[0][2]
    live = sum(p.numel() for n_, p in Toy().named_parameters() if not n_.startswith('dead'))
    assert st['elements'] == live and st['buckets'] >= 2
    assert st['launched_from_hooks'] == 2 * st['buckets'], st      # steps 1 and 2: every bucket left from a hook
    assert st['launched_at_finish'] == st['buckets'], st           # step 0 (calibration) only


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_allreduce_is_identity():
    m = Toy()
    m(torch.randn(4, 6)).sum().backward()
    before = [None if p.grad is None else p.grad.clone() for p in m.parameters()]
    D.allreduce_gradients(list(m.parameters()), world=1)
    for b, p in zip(before, m.parameters()):
        assert (b is None and p.grad is None) or torch.equal(b, p.grad)
