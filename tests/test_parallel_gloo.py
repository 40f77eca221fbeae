"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding + the two-collective gradient
exchange of xlxmert_b200.parallel (flat encoder arena + flat remainder, unused parameters skipped)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xlxmert_b200.config import TINY_DIMS
    from xlxmert_b200.parallel import allreduce_gradients, shard_batch
    from xlxmert_b200.pretraining import B200XLxmertForPretraining
    torch.manual_seed(0)
    m = B200XLxmertForPretraining(TINY_DIMS, num_clusters=TINY_DIMS.num_clusters)
    enc = m.bert.encoder
    # emulate what the encoder backward leaves behind: gradients that are views into one flat arena
    names = [n for n, _ in enc.named_parameters()]
    params = dict(enc.named_parameters())
    total = sum(p.numel() for p in params.values())
    arena = torch.full((total,), float(rank + 1))
    off = 0
    for n in names:
        p = params[n]
        p.grad = arena[off:off + p.numel()].view_as(p)
        off += p.numel()
    enc.last_grad_arena = arena
    # non-encoder gradients: some present, some absent (task-dependent unused parameters)
    m.mask_feat.grad = torch.full_like(m.mask_feat, 10.0 * (rank + 1))
    m.bert.pooler.dense.weight.grad = torch.full_like(m.bert.pooler.dense.weight, 3.0 * (rank + 1))
    calls = allreduce_gradients(m)
    ok = calls == 2
    mean = (1 + world) / 2.0
    ok &= bool(torch.allclose(arena, torch.full_like(arena, mean)))
    ok &= bool(torch.allclose(params[names[3]].grad, torch.full_like(params[names[3]], mean)))   # still views
    ok &= params[names[3]].grad.untyped_storage().data_ptr() == arena.untyped_storage().data_ptr()
    ok &= bool(torch.allclose(m.mask_feat.grad, torch.full_like(m.mask_feat, 10.0 * mean)))
    ok &= bool(torch.allclose(m.bert.pooler.dense.weight.grad, torch.full_like(m.bert.pooler.dense.weight, 3.0 * mean)))
    ok &= m.cls.predictions.bias.grad is None                                                      # skipped, not zero-filled
    batch = {"input_ids": torch.arange(10).view(5, 2), "sent": ["a"] * 5}
    sh = shard_batch(batch, rank, world)
    ok &= sh["input_ids"].shape[0] == 2 and int(sh["input_ids"][0, 0]) == rank * 4 and sh["sent"] == ["a"] * 5
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
