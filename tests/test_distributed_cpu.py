"""world_size-2 gloo test of the multi-GPU host logic: edits are sharded by the reference's rule (inference.py:126-128),
there is no data-path collective, timing is the max over ranks, and the aggregate is units / max time (bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from loongx_b200.sampler import shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(n_items, rank, world)
    mine = torch.zeros(n_items, dtype=torch.int64)
    mine[s:e] = 1  # "process" my shard; no exchange of edit data between ranks
    elapsed = torch.tensor([0.25 * (rank + 1)], dtype=torch.float64)  # pretend per-rank device time
    dist.barrier()
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    cover = mine.clone()
    dist.all_reduce(cover, op=dist.ReduceOp.SUM)  # test-only check that shards tile the work exactly once
    if rank == 0:
        q.put((cover.tolist(), float(elapsed.item()), n_items / float(elapsed.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_two_rank_sharding_and_max_timing(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    cover, t_max, agg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert cover == [1] * n_items
    assert t_max == 0.5 and agg == n_items / 0.5


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from loongx_b200.train import allreduce_mean_

    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, generator=g)  # this rank's flat LoRA-gradient bucket
    mine = flat.clone()
    out = allreduce_mean_(flat)
    assert out.data_ptr() == flat.data_ptr()  # in place: the per-factor views of the bucket see the reduced values
    other = torch.randn(1000, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ok = torch.allclose(flat, (mine + other) / 2, atol=1e-6)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = torch.equal(gathered[0], gathered[1])  # replicas stay identical after the step
    # the bucket layout of a step with CS3 / DGF gradients: [LoRA factors | encoder parameters (complex ones as float
    # pairs)]; ONE all-reduce covers both halves and every parameter receives its slice of the reduced bucket
    import types

    from loongx_b200 import cs3_bwd as CB
    from loongx_b200.train import EncoderBackward, _hand_out

    fA, fB = torch.nn.Parameter(torch.zeros(4, 50)), torch.nn.Parameter(torch.zeros(100, 4))
    enc_params = [torch.nn.Parameter(torch.zeros(3, 7)), torch.nn.Parameter(torch.zeros(5, 6, dtype=torch.complex64)),
                  torch.nn.Parameter(torch.zeros(11))]
    n_lora, n_enc = 600, CB.grad_elements(enc_params)
    bucket = torch.randn(n_lora + n_enc, generator=torch.Generator().manual_seed(200 + rank))
    tr = types.SimpleNamespace(factors={"m": types.SimpleNamespace(A=fA, B=fB)}, n_lora_grad=n_lora, grad_flat=bucket,
                               grad_extra=bucket[n_lora:])
    enc = EncoderBackward(model=None, ctx=None, params=enc_params, trainer=tr)
    assert n_enc == 24 + 60 + 12 and sum(v.numel() for v in enc.views.values()) == 21 + 60 + 11  # 16-byte aligned slices
    allreduce_mean_(tr.grad_flat)
    both = [torch.randn(n_lora + n_enc, generator=torch.Generator().manual_seed(200 + r)) for r in range(world)]
    want = (both[0] + both[1]) / 2
    got = _hand_out(tr, tr.grad_flat, enc)
    flat_again = torch.cat([(torch.view_as_real(t) if t.is_complex() else t).reshape(-1) for t in got])
    lay = CB.grad_layout(enc_params)[0]
    want = torch.cat([want[:n_lora]] + [want[n_lora + o:n_lora + o + n] for o, n in lay])
    ok = ok and torch.allclose(flat_again, want, atol=1e-6) and [tuple(t.shape) for t in got] == [(4, 50), (100, 4), (3, 7), (5, 6), (11,)] \
        and got[3].is_complex()
    if rank == 0:
        q.put((bool(ok), bool(same)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_mean():
    """The one real collective of the path: DDP's mean all-reduce of the trainable gradients, here
    one call over the flat LoRA-gradient bucket (gloo stands in for NCCL on the CPU)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok and same
    # single process / no process group: identity
    from loongx_b200.train import allreduce_mean_

    x = torch.arange(4.0)
    assert torch.equal(allreduce_mean_(x.clone()), x)
