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
        self._in: List[Tuple[torch.Tensor, torch.Tensor]] = []      # two (x, task_id) device buffers
        self._in_free = [torch.cuda.Event(), torch.cuda.Event()]   # buffer may be overwritten (its forward has consumed it)
        self._shape = None

    def _buffers(self, x: torch.Tensor, tid: torch.Tensor):
        key = (tuple(x.shape), x.dtype, tuple(tid.shape), tid.dtype)
        if key != self._shape:
            self._in = [(torch.empty(x.shape, dtype=x.dtype, device=self.device), torch.empty(tid.shape, dtype=tid.dtype, device=self.device))
                        for _ in range(2)]
            self._shape = key
        return self._in

    @torch.no_grad()
    def restore_stream(self, cubes: Iterable[Tuple[torch.Tensor, torch.Tensor]], outs: Sequence[torch.Tensor]) -> None:
        """cubes: (x_host [B,C,H,W] pinned, task_id_host) pairs; outs[i]: pinned host tensor receiving the restored cube i.
        Returns once every copy has been *enqueued*; synchronise the device (or ``self.d2h``) before reading ``outs``."""
        compute = torch.cuda.current_stream(self.device)
        staged = None  # (slot, event: inputs landed)
        it = iter(cubes)

        def stage(slot, item):
            x_h, t_h = item
            xd, td = self._buffers(x_h, t_h)[slot]
            with torch.cuda.stream(self.h2d):
                self.h2d.wait_event(self._in_free[slot])
                xd.copy_(x_h, non_blocking=True)
                td.copy_(t_h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.h2d)
            return slot, ev

        for ev in self._in_free:
            ev.record(compute)
        first = next(it, None)
        if first is None:
            return
        staged = stage(0, first)
        i = 0
        while staged is not None:
            slot, landed = staged
            nxt = next(it, None)
            staged = stage(slot ^ 1, nxt) if nxt is not None else None   # cube i+1 travels while cube i computes
            compute.wait_event(landed)
            xd, td = self._in[slot]
            y = self.net(xd, td)
            self._in_free[slot].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                outs[i].copy_(y, non_blocking=True)
            y.record_stream(self.d2h)
            i += 1

