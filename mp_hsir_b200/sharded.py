"""Exact row-sharded inference of ONE scene over G GPUs (BASELINE config 3).

``test.py:170`` feeds the whole 31x512x512 cube to one GPU.  Independent tiles are not parity-safe: the global spectral
attention normalises and correlates over ALL H*W pixels (net/MP_HSIR.py:104-110) and the shifted windows roll cyclically
over the whole image (:672, :694).  Here every rank owns a band of image rows and the result equals the single-GPU
forward up to the summation order of one small all-reduce per block:

  * layout: per resolution level a rank holds its ``Hb = H/G`` rows plus an 8-row halo above and below (one window row),
    token-major, so a halo is one contiguous row range.  Local row i is scene row ``(rank*Hb - 8 + i) mod H``.
  * halo refresh (``Band.exchange``): before every PGSSTB, every dense 3x3 conv and every PromptFusion the halos of the
    input are copied from the neighbour ranks' own rows — cyclically, the last rank's bottom halo is the first rank's top
    rows, which is exactly what ``torch.roll`` wraps.  NCCL send/recv between neighbours, 8*W*C floats each way.
  * window attention, LayerNorm, all 1x1 GEMMs, the gated MLP and the local spectral gate are purely local on the
    ``Hb+16`` rows (windows that wrap inside the band lie entirely in the halo and are recomputed by the neighbour); the
    Swin mask is evaluated in scene coordinates (``mphsir_window_attn_band_fwd``).
  * 3x3 convs / depthwise convs run on the rows ``[a, b)`` a rank may touch: the whole local image on inner ranks, but
    cut at the scene's top (rank 0) / bottom (last rank), so that the kernels' zero padding is the conv's padding.
  * global spectral attention / PromptFusion MDTA: ``q^T k``, ``sum q^2``, ``sum k^2`` over the rank's OWN rows
    (``mphsir_dwgram_band_fwd``) -> ``mphsir_gram_reduce`` -> ONE all-reduce (sum) of heads*(c*c+2c) floats (2 176 at
    level 1, <= 18.8k for the remote-sensing model) -> softmax / fold on every rank.  24 per forward.
  * TVSP depends on (task ids, shape) only: every rank computes the full-resolution prompt once (cached) and keeps its rows.

Communication back-ends: ``PeerComm`` (one process per GPU of one node; each collective is ONE libmphsir kernel over NVLink
peer memory — CUDA IPC windows, device-side sequence numbers, CUDA-graph capturable; mp_hsir_b200/csrc/peer.cu),
``NcclComm`` (torch.distributed send/recv + all_reduce) and ``ThreadComm`` (G virtual ranks as threads on ONE GPU — the same
code path with device-local copies; used by the single-GPU parity tests and as a debugging aid).  Host plumbing only: all
arithmetic is libmphsir launches.
"""
from __future__ import annotations

import threading
from typing import List, Optional

import torch

from . import lib
from .engine import Engine, Workspace
from .lib import View

HALO = 8  # rows: one window row, covers the 4-row roll and every 3x3 stencil


def band_rows(H: int, rank: int, world: int, scale: int = 1):
    """[r0, r1) scene rows owned by `rank` at resolution level `scale` (1, 2, 4).  H must split into `world` bands of
    whole level-3 windows: H % (32 * world) == 0."""
    if H % (32 * world):
        raise ValueError(f"a {H}-row scene cannot be split into {world} bands of whole windows at every level "
                         f"(H must be a multiple of {32 * world})")
    hb = H // scale // world
    return rank * hb, (rank + 1) * hb


class Band:
    """Geometry of one rank's band at one resolution level + the halo exchange."""

    def __init__(self, comm, Hg: int, W: int):
        self.comm, self.Hg, self.W = comm, Hg, W
        self.halo = HALO
        self.Hb = Hg // comm.world
        if self.Hb * comm.world != Hg or self.Hb % 8:
            raise ValueError(f"{Hg} rows do not split into {comm.world} bands of whole windows")
        self.Hloc = self.Hb + 2 * HALO
        self.r0 = comm.rank * self.Hb
        self.y0 = (self.r0 - HALO) % Hg                       # scene row of local row 0
        self.a = HALO if comm.rank == 0 else 0                # rows [a, b) exist in the scene without wrapping
        self.b = self.Hb + HALO if comm.rank == comm.world - 1 else self.Hloc
        self.N = self.Hloc * W

    def exchange(self, v: View) -> None:
        """top halo <- previous rank's last own rows, bottom halo <- next rank's first own rows (cyclic).  `v` is a
        [Hloc*W, cols] view of a workspace matrix; whole rows of the underlying matrix travel (all its columns)."""
        base = v.keep.reshape(-1)
        off = (v.ptr - v.keep.data_ptr()) // 4
        row0 = off - off % v.ld                               # start of local row 0 in the underlying matrix
        n = HALO * self.W * v.ld
        rows = lambda r: row0 + r * self.W * v.ld             # noqa: E731
        self.comm.halo(base, top=rows(0), own_first=rows(HALO), own_last=rows(self.Hb), bottom=rows(self.Hb + HALO), n=n)


# ------------------------------------------------------------------------------------------------
# communication back-ends
# ------------------------------------------------------------------------------------------------


class NcclComm:
    """one process per GPU (torch.distributed, any backend with send/recv + all_reduce: nccl on GPUs)"""

    name = "torch.distributed send/recv (halos) + all_reduce (Gram statistics)"

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.halo_exchanges = self.all_reduces = 0

    def halo(self, base: torch.Tensor, top: int, own_first: int, own_last: int, bottom: int, n: int) -> None:
        self.halo_exchanges += 1
        if self.world == 1:
            base[top:top + n].copy_(base[own_last:own_last + n])
            base[bottom:bottom + n].copy_(base[own_first:own_first + n])
            return
        d = self.dist
        prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        # with two ranks prev == nxt: sends and receives to one peer pair up in issue order, so the first receive
        # (peer's first rows) must be the BOTTOM halo
        ops = [d.P2POp(d.isend, base[own_first:own_first + n], prev, self.group),
               d.P2POp(d.isend, base[own_last:own_last + n], nxt, self.group),
               d.P2POp(d.irecv, base[bottom:bottom + n], nxt, self.group),
               d.P2POp(d.irecv, base[top:top + n], prev, self.group)]
        for req in d.batch_isend_irecv(ops):
            req.wait()

    def all_reduce(self, t: torch.Tensor) -> None:
        self.all_reduces += 1
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


class PeerComm:
    """one process per GPU of ONE node: both collectives are single libmphsir kernels over NVLink peer memory
    (mp_hsir_b200/csrc/peer.cu) — every rank's window (mailboxes, gather slots, flags) is exported through CUDA IPC and mapped
    by all peers; torch.distributed is used once, to swap the 64-byte handles.  CUDA-graph capturable (the sequence numbers
    live in device memory)."""

    name = "libmphsir peer-memory kernels over NVLink (CUDA IPC windows): 1 kernel per halo exchange / all-reduce"

    def __init__(self, device, halo_cap: int = 1 << 20, ar_cap: int = 1 << 15, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("PeerComm serves the GPUs of one node (<= 8 ranks)")
        self.device = torch.device(device)
        self.halo_cap, self.ar_cap = int(halo_cap), int(ar_cap)
        self.halo_exchanges = self.all_reduces = 0
        L = lib.load()
        with torch.cuda.device(self.device):
            win = C.c_void_p()
            lib._check(L.mphsir_peer_window_alloc(L.mphsir_peer_window_bytes(self.halo_cap, self.ar_cap), C.byref(win)), "peer_window_alloc")
            lib.LAUNCHES -= 1
            self._own = win.value
            handle = C.create_string_buffer(64)
            lib._check(L.mphsir_peer_export(self._own, handle), "peer_export")
            lib.LAUNCHES -= 1
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._mapped = []
            table = (C.c_void_p * self.world)()
            for q in range(self.world):
                if q == self.rank:
                    table[q] = self._own
                    continue
                m = C.c_void_p()
                lib._check(L.mphsir_peer_open(handles[q], C.byref(m)), f"peer_open(rank {q})")
                lib.LAUNCHES -= 1
                self._mapped.append(m.value)
                table[q] = m.value
            self._table = table
            dist.barrier(group=group)        # every window is mapped everywhere before the first kernel touches one

    def halo(self, base: torch.Tensor, top: int, own_first: int, own_last: int, bottom: int, n: int) -> None:
        self.halo_exchanges += 1
        p0 = base.data_ptr()
        L = lib.load()
        lib._launch("peer_halo_exchange",
                    lambda: L.mphsir_peer_halo_exchange(self._table, self.rank, self.world, self.halo_cap, self.ar_cap,
                                                        p0 + 4 * own_first, p0 + 4 * own_last, p0 + 4 * top, p0 + 4 * bottom, n,
                                                        lib.stream_ptr()),
                    lambda: (0.0, 16.0 * n, "peer_halo_exchange"))

    def all_reduce(self, t: torch.Tensor) -> None:
        self.all_reduces += 1
        L = lib.load()
        lib._launch("peer_all_reduce",
                    lambda: L.mphsir_peer_all_reduce(self._table, self.rank, self.world, self.halo_cap, self.ar_cap, t.data_ptr(),
                                                     t.numel(), lib.stream_ptr()),
                    lambda: (0.0, 4.0 * t.numel() * (2 * self.world + 1), "peer_all_reduce"))

    def close(self) -> None:
        L = lib.load()
        torch.cuda.synchronize(self.device)
        for m in self._mapped:
            L.mphsir_peer_close(m)
        self._mapped = []
        if self._own:
            L.mphsir_peer_window_free(self._own)
            self._own = None


class _ThreadWorld:
    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots: List[Optional[tuple]] = [None] * world


class ThreadComm:
    """`world` virtual ranks as threads of one process on ONE GPU.  Every rank posts its buffer, a barrier orders the
    launches on the shared stream, copies are device-local.  The all-reduce sums in rank order on every rank (bit-identical
    results on all ranks, like NCCL's)."""

    name = "in-process virtual ranks (device-local copies)"

    def __init__(self, shared: _ThreadWorld, rank: int):
        self.shared, self.rank, self.world = shared, rank, shared.world
        self.halo_exchanges = self.all_reduces = 0

    def halo(self, base: torch.Tensor, top: int, own_first: int, own_last: int, bottom: int, n: int) -> None:
        self.halo_exchanges += 1
        sh = self.shared
        sh.slots[self.rank] = (base, own_first, own_last)
        sh.barrier.wait()
        pb, _, p_last = sh.slots[(self.rank - 1) % self.world]
        nb, n_first, _ = sh.slots[(self.rank + 1) % self.world]
        base[top:top + n].copy_(pb[p_last:p_last + n])
        base[bottom:bottom + n].copy_(nb[n_first:n_first + n])
        sh.barrier.wait()

    def all_reduce(self, t: torch.Tensor) -> None:
        self.all_reduces += 1
        sh = self.shared
        sh.slots[self.rank] = (t,)
        sh.barrier.wait()
        total = sh.slots[0][0].clone()
        for r in range(1, self.world):
            total += sh.slots[r][0]
        sh.barrier.wait()          # everyone has read every contribution
        t.copy_(total)
        sh.barrier.wait()


# ------------------------------------------------------------------------------------------------
# the sharded forward
# ------------------------------------------------------------------------------------------------


class ShardedEngine(Engine):
    """Forward of one rank's band.  ``forward_band(x_band, task_id, H)``: x_band [1, C, H/G, W] = this rank's rows of the
    scene (rank r owns rows [r*H/G, (r+1)*H/G)); returns the restored band."""

    def __init__(self, net, comm):
        super().__init__(net)
        if self.prec == lib.PREC_FP32_SIMT:
            raise ValueError("sharded scenes run on the tensor-core precisions ('fp32' = bf16x3, or 'bf16')")
        self.comm = comm
        self.ws = Workspace(self.device, zero_new=True)
        self._band_prompt_key = {}
        # replay the band's whole launch sequence — kernels AND the NCCL send/recv / all-reduce calls — as one CUDA graph
        # per (scene shape, task ids): at 4-8 GPUs a band's kernels are a few microseconds each and the eager launches
        # (~290 kernels + ~50 collectives per scene) would otherwise bound the step
        self.use_cuda_graph = False
        self._band_graphs = {}

    # prompts: full-resolution TVSP (cached, input independent) -> this rank's rows of the fusion buffer
    def _band_prompt(self, name: str, clip_b, weights, bd: Band, D: int, dst: View, task_key) -> None:
        full = self.ws.mat(name + ".full", bd.Hg * bd.W, D)
        band, self.band = self.band, None      # TVSP is a whole-image computation on every rank
        try:
            self._tvsp_cached(name, clip_b, weights, 1, bd.Hg, bd.W, full, task_key)
        finally:
            self.band = band
        key = (self._tvsp_valid.get(name), dst.ptr, dst.ld)
        if task_key is not None and self._band_prompt_key.get(name) == key:
            return
        W = bd.W
        dst.torch()[HALO * W:(HALO + bd.Hb) * W].copy_(full.torch()[bd.r0 * W:(bd.r0 + bd.Hb) * W])
        self._band_prompt_key[name] = key

    def _conv_rows(self, X: View, Wt, Y: View, bd_in: Band, r0: int, r1: int, y_row0: int, Cin: int, N: int, mode: int):
        """dense 3x3 conv over local input rows [r0, r1) of level `bd_in`; the result goes to `Y` starting at its local
        row y_row0 (Y may live at another resolution level: pixel (un)shuffle)."""
        W = bd_in.W
        Wy = {lib.CONV_TOKENS: W, lib.CONV_UNSHUFFLE: W // 2, lib.CONV_SHUFFLE: 2 * W}[mode]
        yv = Y.rows_slice(y_row0 * Wy, Y.rows)
        self._conv(X.rows_slice(r0 * W, r1 * W), Wt, yv.ptr, yv.ld, 1, r1 - r0, W, Cin, N, mode)

    @torch.no_grad()
    def forward_band(self, x_band: torch.Tensor, task_id: torch.Tensor, H: int) -> torch.Tensor:
        cfg, comm = self.cfg, self.comm
        if x_band.dim() != 4 or x_band.shape[0] != 1 or x_band.shape[1] != cfg.in_channel:
            raise ValueError(f"expected one band [1,{cfg.in_channel},H/G,W], got {tuple(x_band.shape)}")
        W = x_band.shape[3]
        r0, r1 = band_rows(H, comm.rank, comm.world)
        if x_band.shape[2] != r1 - r0 or W % 32:
            raise ValueError(f"rank {comm.rank}/{comm.world} owns rows [{r0},{r1}) of a {H}x{W} scene (W % 32 == 0), got {tuple(x_band.shape)}")
        if x_band.device != self.device:
            raise RuntimeError(f"input on {x_band.device} but parameters on {self.device}")
        self._ensure_packed()
        task_key = (tuple(task_id.shape), tuple(task_id.reshape(-1).tolist())) if self.cache_prompts else None
        x = x_band.detach().to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            if self.use_cuda_graph and task_key is not None and lib.PROFILER is None:
                return self._run_band_graphed(x, task_id, H, task_key)
            return self._run_band(x, self.task_weights(task_id), H, task_key)

    def _run_band_graphed(self, x: torch.Tensor, task_id: torch.Tensor, H: int, task_key) -> torch.Tensor:
        """eager on the first two calls of a (shape, task ids) key (allocations, prompt cache, NCCL channel set-up), then
        capture once and replay.  Every rank must take the same path at the same call, which holds when all ranks run the
        same sequence of scenes (they do: one scene is one collective operation)."""
        key = (tuple(x.shape), H, task_key, self._pack_serial)
        ent = self._band_graphs.get(key)
        if ent is not None and ent[0] == "graph" and ent[1] != self.ws.generation:
            ent = None
        if ent is None or ent[0] == "warm":
            n = 0 if ent is None else ent[1]
            if n < 2:
                self._band_graphs[key] = ("warm", n + 1)
                return self._run_band(x, self.task_weights(task_id), H, task_key)
            sx = x.clone()
            sw = self.task_weights(task_id).clone()
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = lib.LAUNCHES
            with torch.cuda.graph(g):
                so = self._run_band(sx, sw, H, task_key)
            ent = ("graph", self.ws.generation, g, sx, so, lib.LAUNCHES - n0)
            lib.LAUNCHES = n0
            self._band_graphs[key] = ent
        _, _, g, sx, so, n_kernels = ent
        sx.copy_(x)
        g.replay()
        lib.LAUNCHES += n_kernels
        return so.clone()

    def _run_band(self, x: torch.Tensor, weights: torch.Tensor, H: int, task_key) -> torch.Tensor:
        cfg, P, ws, comm = self.cfg, self.packed, self.ws, self.comm
        W = x.shape[3]
        d = cfg.dim
        b1, b2, b3 = Band(comm, H, W), Band(comm, H // 2, W // 2), Band(comm, H // 4, W // 4)
        W2, W3 = W // 2, W // 4
        T = cfg.task_classes
        try:
            clip_b = ws.flat("clip_b", 512)
            lib.text_prompt(weights, P["clip"], clip_b, 1, T)

            # ---- level 1 ------------------------------------------------------------------------
            self.band = b1
            tok = ws.mat("tok_in", b1.N, P["cin_p"])
            lib.nchw_to_tokens(x, tok.rows_slice(HALO * W, (HALO + b1.Hb) * W))
            b1.exchange(tok)
            x1 = ws.mat("x1", b1.N, d)
            self._conv_rows(tok, P["patch_embed"], x1, b1, b1.a, b1.b, b1.a, P["cin_p"], d, lib.CONV_TOKENS)
            fcat1 = ws.mat("fcat1", b1.N, 2 * d)          # [e1 | prompt1]
            e1 = fcat1.cols_slice(0, d)
            self._stage("encoder_level1", x1, e1, 1, b1.Hloc, W)
            self._band_prompt("prompt1", clip_b, weights, b1, d, fcat1.cols_slice(d, 2 * d), task_key)
            b1.exchange(fcat1)                            # e1 and prompt1 halos in one message

            # ---- level 2 ------------------------------------------------------------------------
            # level-1 local row i is level-2 local row (i + 8) / 2
            x2 = ws.mat("x2", b2.N, 2 * d)
            self._conv_rows(e1, P["down1_2"], x2, b1, b1.a, b1.b, (b1.a + HALO) // 2, d, d // 2, lib.CONV_UNSHUFFLE)
            self.band = b2
            fcat2 = ws.mat("fcat2", b2.N, 4 * d)          # [e2 | prompt2]
            e2 = fcat2.cols_slice(0, 2 * d)
            self._stage("encoder_level2", x2, e2, 1, b2.Hloc, W2)
            self._band_prompt("prompt2", clip_b, weights, b2, 2 * d, fcat2.cols_slice(2 * d, 4 * d), task_key)
            b2.exchange(fcat2)

            # ---- level 3 ------------------------------------------------------------------------
            x3 = ws.mat("x3", b3.N, 4 * d)
            self._conv_rows(e2, P["down2_3"], x3, b2, b2.a, b2.b, (b2.a + HALO) // 2, 2 * d, d, lib.CONV_UNSHUFFLE)
            self.band = b3
            lat = ws.mat("lat", b3.N, 4 * d)
            self._stage("latent", x3, lat, 1, b3.Hloc, W3)
            b3.exchange(lat)

            # ---- back to level 2: level-3 local row k is level-2 local rows 2k - 8, 2k - 7 ---------
            cat2 = ws.mat("cat2", b2.N, 4 * d)            # [up3_2(latent) | fusion2]
            k0, k1 = max(b3.a, HALO // 2), min(b3.b, b3.Hb + HALO + HALO // 2)
            self._conv_rows(lat, P["up3_2"], cat2, b3, k0, k1, 2 * k0 - HALO, 4 * d, 8 * d, lib.CONV_SHUFFLE)
            self.band = b2
            self._fusion("fusion2", fcat2, cat2.cols_slice(2 * d, 4 * d), 1, b2.Hloc, W2)
            d2in = ws.mat("d2in", b2.N, 2 * d)
            self._gemm(cat2, P["reduce_chan_level2"], d2in, 2 * d)
            d2 = ws.mat("d2", b2.N, 2 * d)
            self._stage("decoder_level2", d2in, d2, 1, b2.Hloc, W2)
            b2.exchange(d2)

            # ---- back to level 1 ------------------------------------------------------------------
            cat1 = ws.mat("cat1", b1.N, 2 * d)            # [up2_1(d2) | fusion1]
            k0, k1 = max(b2.a, HALO // 2), min(b2.b, b2.Hb + HALO + HALO // 2)
            self._conv_rows(d2, P["up2_1"], cat1, b2, k0, k1, 2 * k0 - HALO, 2 * d, 4 * d, lib.CONV_SHUFFLE)
            self.band = b1
            self._fusion("fusion1", fcat1, cat1.cols_slice(d, 2 * d), 1, b1.Hloc, W)
            dd1 = ws.mat("dd1", b1.N, 2 * d)
            self._stage("decoder_level1", cat1, dd1, 1, b1.Hloc, W)
            ref = ws.mat("ref", b1.N, 2 * d)
            self._stage("refinement", dd1, ref, 1, b1.Hloc, W)
            b1.exchange(ref)

            # output conv + global residual (:841) over the rows [a, b); the band is cut out of the NCHW result
            Hv, o0 = b1.b - b1.a, HALO - b1.a
            res = ws.flat("out_res", cfg.in_channel * Hv * W).view(-1)[: cfg.in_channel * Hv * W].view(1, cfg.in_channel, Hv, W)
            res[:, :, o0:o0 + b1.Hb].copy_(x)
            out = ws.flat("out_ext", cfg.out_channel * Hv * W).view(-1)[: cfg.out_channel * Hv * W].view(1, cfg.out_channel, Hv, W)
            rv = ref.rows_slice(b1.a * W, b1.b * W)
            self._conv(rv, P["output"], out.data_ptr(), 0, 1, Hv, W, 2 * d, cfg.out_channel, lib.CONV_NCHW_RES, R=res)
            return out[:, :, o0:o0 + b1.Hb].contiguous()
        finally:
            self.band = None


def restore_scene_virtual(net, x: torch.Tensor, task_id: torch.Tensor, world: int) -> torch.Tensor:
    """The sharded forward of `world` ranks run as threads on ONE GPU (ThreadComm): returns the assembled scene.  Same code
    path as the multi-GPU run except for the transport; used by the parity tests and for debugging."""
    H = x.shape[2]
    shared = _ThreadWorld(world)
    engines = [ShardedEngine(net, ThreadComm(shared, r)) for r in range(world)]
    engines[0]._ensure_packed()
    for e in engines[1:]:   # weights are read-only: share the packed images
        e.packed, e._versions, e._pack_serial = engines[0].packed, engines[0]._versions, engines[0]._pack_serial
    outs: List[Optional[torch.Tensor]] = [None] * world
    errors: List[BaseException] = []
    stream = torch.cuda.current_stream(x.device)

    def run(r: int):
        try:
            with torch.cuda.device(x.device), torch.cuda.stream(stream):
                r0, r1 = band_rows(H, r, world)
                outs[r] = engines[r].forward_band(x[:, :, r0:r1].contiguous(), task_id, H)
        except BaseException as e:  # noqa: BLE001 - surface the failure in the caller and release the other ranks
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return torch.cat(outs, dim=2)


def restore_scene(net, x_band: torch.Tensor, task_id: torch.Tensor, H: int, engine: Optional[ShardedEngine] = None):
    """One process per GPU: restore this rank's band of an H-row scene (torch.distributed must be initialised).  Pass the
    returned engine back in to keep its workspace, packed weights and prompt cache."""
    if engine is None:
        engine = ShardedEngine(net, NcclComm())
    return engine.forward_band(x_band, task_id, H), engine
