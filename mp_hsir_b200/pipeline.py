"""Host-to-host restoration of a stream of cubes with the copies hidden behind the compute.

``test.py`` of the reference restores one cube after another (test.py:157-188: host tensor -> ``.to(device)`` -> ``net`` ->
``.cpu()``), paying the PCIe transfers of cube i (32.5 MB each way for 31x512x512) in series with its forward.
``restore_stream`` keeps the same per-cube call (``net(x, task_id)`` on the device) but runs the host->device copy of
cube i+1 and the device->host copy of cube i-1 on two copy streams while cube i computes: double-buffered device inputs,
CUDA events for the hand-offs, ``record_stream`` on the outputs.  Plumbing only — no arithmetic here.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch


class HostPipeline:
    def __init__(self, net, device=None):
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self._in: List[Tuple[torch.Tensor, torch.Tensor]] = [None, None]   # two (x, task_id) device buffers, keyed per slot
        self._in_free = [torch.cuda.Event(), torch.cuda.Event()]   # buffer may be overwritten (its forward has consumed it)
        self._shape = [None, None]

    def _buffer(self, slot: int, x: torch.Tensor, tid: torch.Tensor):
        """device staging buffers of ONE slot; a cube of another shape (test sets mix sizes) re-allocates only this slot —
        the other slot may hold a cube that is staged but not yet consumed"""
        key = (tuple(x.shape), x.dtype, tuple(tid.shape), tid.dtype)
        if key != self._shape[slot]:
            self._in[slot] = (torch.empty(x.shape, dtype=x.dtype, device=self.device),
                              torch.empty(tid.shape, dtype=tid.dtype, device=self.device))
            self._shape[slot] = key
        return self._in[slot]

    @torch.no_grad()
    def restore_stream(self, cubes: Iterable[Tuple[torch.Tensor, torch.Tensor]], outs: Sequence[torch.Tensor]) -> None:
        """cubes: (x_host [B,C,H,W] pinned, task_id_host) pairs; outs[i]: pinned host tensor receiving the restored cube i.
        Returns once every copy has been *enqueued*; synchronise the device (or ``self.d2h``) before reading ``outs``."""
        compute = torch.cuda.current_stream(self.device)
        staged = None  # (slot, event: inputs landed, device input, device task ids, host task ids)
        it = iter(cubes)

        def stage(slot, item):
            x_h, t_h = item
            with torch.cuda.stream(self.h2d):
                self.h2d.wait_event(self._in_free[slot])   # before a possible re-allocation, too: the old buffer's forward is done
                xd, td = self._buffer(slot, x_h, t_h)
                xd.copy_(x_h, non_blocking=True)
                td.copy_(t_h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.h2d)
            return slot, ev, xd, td, t_h

        for ev in self._in_free:
            ev.record(compute)
        first = next(it, None)
        if first is None:
            return
        staged = stage(0, first)
        i = 0
        while staged is not None:
            slot, landed, xd, td, t_host = staged
            nxt = next(it, None)
            staged = stage(slot ^ 1, nxt) if nxt is not None else None   # cube i+1 travels while cube i computes
            compute.wait_event(landed)
            # the HOST copy of the task ids goes to the module (it keys the prompt cache; no synchronising read-back)
            y = self.net(xd, t_host)
            self._in_free[slot].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                outs[i].copy_(y, non_blocking=True)
            y.record_stream(self.d2h)
            i += 1

