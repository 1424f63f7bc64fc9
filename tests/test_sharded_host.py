"""CPU: partition math and communication of the row-sharded single-scene path (mp_hsir_b200/sharded.py) — band geometry,
halo exchange and Gram all-reduce over gloo with world_size 2 and 3 (2 is the prev == next special case), and the
in-process ThreadComm used by the single-GPU parity tests."""
import os
import socket
import threading

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mp_hsir_b200.lib import View
from mp_hsir_b200.sharded import HALO, Band, NcclComm, ThreadComm, _ThreadWorld, band_rows


class _FakeComm:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world


def test_band_rows_partition_every_level():
    for H, world in ((512, 1), (512, 2), (512, 4), (512, 8), (256, 2), (128, 4)):
        for scale in (1, 2, 4):
            spans = [band_rows(H, r, world, scale) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == H // scale
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all((e - s) % 8 == 0 for s, e in spans)          # whole windows at every level
    with pytest.raises(ValueError):
        band_rows(512, 0, 3)                                          # 512 rows do not split into 3 bands of whole windows
    with pytest.raises(ValueError):
        band_rows(96, 0, 2)                                           # level-3 band would be 12 rows


def test_band_geometry_scene_coordinates():
    H, W, world = 256, 64, 4
    for r in range(world):
        b = Band(_FakeComm(r, world), H, W)
        assert b.Hb == 64 and b.Hloc == 80 and b.N == 80 * W
        assert b.y0 == (r * 64 - HALO) % H                          # scene row of local row 0 (cyclic)
        # rows [a, b) are the local rows that exist in the scene without wrapping: cut at the scene edge on the outer ranks
        assert (b.a, b.b) == (HALO if r == 0 else 0, 72 if r == world - 1 else 80)
        # the scene's wrap window (shifted rows H-8..H-1 = scene rows H-4..H-1, 0..3) is the first shifted window row of rank 0 ...
        if r == 0:
            assert [(ys + b.y0) % H for ys in range(8)] == list(range(H - 8, H))
    one = Band(_FakeComm(0, 1), 64, 64)
    assert (one.a, one.b, one.y0) == (HALO, 72, 56)


def _exchange_case(comm, Hb=16, W=8, ld=12):
    """a [Hloc*W, ld] matrix whose own rows hold rank*1000 + scene row; after the exchange the halos hold the neighbours'"""
    bd = Band(comm, Hb * comm.world, W)
    t = torch.full((bd.Hloc * W, ld), -1.0)
    own = torch.arange(bd.r0, bd.r0 + Hb, dtype=torch.float32).repeat_interleave(W)
    t[HALO * W:(HALO + Hb) * W] = (own + 1000.0 * comm.rank).unsqueeze(1)
    v = View(t.data_ptr() + 4 * 4, ld, bd.Hloc * W, 8, t)          # a column slice: whole rows of the matrix travel
    bd.exchange(v)
    Hg = bd.Hg
    top = t[:HALO * W, 0].view(HALO, W)[:, 0]
    bot = t[(HALO + Hb) * W:, ld - 1].view(HALO, W)[:, 0]
    prev, nxt = (comm.rank - 1) % comm.world, (comm.rank + 1) % comm.world
    exp_top = torch.tensor([(bd.r0 - HALO + i) % Hg + 1000.0 * prev for i in range(HALO)])
    exp_bot = torch.tensor([(bd.r0 + Hb + i) % Hg + 1000.0 * nxt for i in range(HALO)])
    g = torch.full((5,), float(comm.rank + 1))
    comm.all_reduce(g)
    return bool(torch.equal(top, exp_top) and torch.equal(bot, exp_bot)), float(g[0])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok, s = _exchange_case(NcclComm())
        dist.barrier()
        out.put((rank, ok, s))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_all_reduce_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(s == world * (world + 1) / 2 for _, _, s in res)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_thread_comm_matches_the_distributed_semantics(world):
    shared = _ThreadWorld(world)
    res = [None] * world

    def run(r):
        res[r] = _exchange_case(ThreadComm(shared, r))

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(ok for ok, _ in res) and all(s == world * (world + 1) / 2 for _, s in res)
