"""``torch.autograd`` visibility for the drop-in module, so that the reference's own training code runs unchanged:

    restored = net(degrad_patch, prompt)            # train.py:58
    loss = l1(torch.clamp(restored, 0, 1), clean)   # train.py:59-61
    loss.backward()                                 # Lightning / GradScaler
    torch.optim.AdamW(net.parameters()).step()      # train.py:69
    DistributedDataParallel(net)                    # train.py:118 (strategy "auto" with >1 device)

One ``autograd.Function`` spans the whole network: its forward is ``TrainEngine.forward_train`` (libmphsir launches,
activations the backward needs stay in the engine's workspace), its backward is ``TrainEngine.backward`` (the hand-written
backward kernels), and it hands one gradient per live ``nn.Parameter`` back to autograd — so ``AccumulateGrad`` runs for
every parameter, which is what DDP's reducer hooks and ``param.grad``-reading optimisers need.  The 8 parameters the
reference never uses (``prompt{1,2}.{text,clip}_linear``) get ``None``, exactly like the reference
(``grad is None`` there too; DDP needs ``find_unused_parameters=True`` for them, as train.py:120's comment says).

``net.trainer().train_step`` stays the fast path (flat buffers, fused AdamW, CUDA-graph replay); this wrapper costs one
extra copy of the flat gradient buffer per step.  Host plumbing only — no arithmetic here.
"""
from __future__ import annotations

import torch


class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, task_id, engine, names, *params):  # noqa: D401 - autograd API
        out, F = engine.autograd_forward(inp, task_id)
        ctx.engine, ctx.F, ctx.names = engine, F, names
        ctx.serial = engine.autograd_serial
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("mp_hsir_b200: no gradient with respect to the input image (the training path of the "
                                      "reference never asks for one, train.py:50-67)")
        eng = ctx.engine
        if d_out is None:
            return (None,) * (4 + len(ctx.names))
        if ctx.serial != eng.autograd_serial:
            raise RuntimeError("mp_hsir_b200: another forward ran on this module since the one being differentiated; saved "
                               "activations live in the engine's workspace, so only the LATEST forward can be back-propagated")
        grads = eng.autograd_backward(ctx.F, d_out)
        ctx.F = None
        return (None, None, None, None) + tuple(grads[n] for n in ctx.names)


def apply(net, inp: torch.Tensor, task_id: torch.Tensor) -> torch.Tensor:
    """``net(inp, task_id)`` with a grad_fn (called by ``MP_HSIR_Net.forward`` when gradients are being recorded)."""
    eng = net.trainer()
    eng.release_param_grads()
    live = [(n, p) for n, p in net.named_parameters() if n in eng.g and p.requires_grad]
    names = tuple(n for n, _ in live)
    return _NetFunction.apply(inp, task_id, eng, names, *[p for _, p in live])
