"""CPU, world_size 2 over gloo: the N>1 path of the hot path is image sharding + one all-gather of fixed-shape outputs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hipie_b200.parallel import all_gather_outputs, shard_indices


def test_shard_indices_match_inference_sampler_convention():
    # total 10 over 4 ranks -> 3,3,2,2 contiguous
    got = [list(shard_indices(10, 4, r)) for r in range(4)]
    assert got == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert [len(shard_indices(64, 8, r)) for r in range(8)] == [8] * 8
    assert sum(len(shard_indices(7, 8, r)) for r in range(8)) == 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = list(shard_indices(total, world, rank))
    g = torch.Generator().manual_seed(0)
    full_logits = torch.randn(total, 5, 7, generator=g)
    full_boxes = torch.rand(total, 5, 4, generator=g)
    out = {"pred_logits": full_logits[idx], "pred_boxes": full_boxes[idx]}
    gathered = all_gather_outputs(out, ["pred_logits", "pred_boxes"])
    ok = torch.equal(gathered["pred_logits"], full_logits) and torch.equal(gathered["pred_boxes"], full_boxes)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_two_rank_gather_restores_global_order(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _worker_packed(rank, world, port, q):
    from hipie_b200.parallel import PackedAllGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    full = {"pred_logits": torch.randn(6, 5, 7, generator=g), "pred_boxes": torch.rand(6, 5, 4, generator=g), "pred_boxious": torch.randn(6, 5, 1, generator=g)}
    idx = list(shard_indices(6, world, rank))
    gather = PackedAllGather()
    ok = True
    for step in range(2):                      # second call reuses the preallocated buffers
        out = {k: v[idx] + step for k, v in full.items()}
        got = gather(out, list(full))
        ok &= all(torch.equal(got[k], full[k] + step) for k in full)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_packed_all_gather_one_collective_restores_global_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
