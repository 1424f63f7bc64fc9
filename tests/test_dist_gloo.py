"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the cube-sharded path (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mp_hsir_b200.parallel import all_reduce_gradients, gather_psnr, max_over_ranks, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(7, rank, world)
        slowest = max_over_ranks(10.0 + 5.0 * rank)
        mean = gather_psnr([30.0 + u for u in range(lo, hi)])
        # DDP step of the trainer: every rank holds the gradient of ITS batch shard in one flat buffer
        flat = torch.full((1000,), float(rank + 1))
        scale = all_reduce_gradients(flat)
        dist.barrier()
        out.put((rank, lo, hi, slowest, mean, float((flat * scale).mean()), float(flat.min()), float(flat.max())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 7)]            # 7 cubes over 2 ranks, disjoint + complete
    assert all(abs(r[3] - 15.0) < 1e-12 for r in res)                  # max over ranks
    assert all(abs(r[4] - (30.0 + 3.0)) < 1e-12 for r in res)          # mean over all 7 cubes
    assert all(r[5] == 1.5 and r[6] == r[7] == 3.0 for r in res)       # summed gradient, mean = sum / world on every rank


def test_shard_range_covers_everything():
    for n in (0, 1, 5, 8, 50):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
