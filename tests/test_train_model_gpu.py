"""GPU: whole-network backward of the CUDA path against the oracle's autograd (itself pinned to the unmodified
reference's gradients by tests/test_train_oracle.py), the clamp+L1 train step, and AdamW bookkeeping."""
import pytest
import torch

from mp_hsir_b200 import MP_HSIR_Net
from mp_hsir_b200.config import NetConfig
from mp_hsir_b200.synth import fill_state_dict_, synthetic_clip_prompt, synthetic_input
from oracle import mp_hsir_oracle as O
from tests.conftest import rel_err
from tests.train_helpers import golden_grads, keep_multipliers, objective_weights, oracle_grads

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(precision):
    cfg = NetConfig.natural()
    net = MP_HSIR_Net(cfg.in_channel, cfg.out_channel, cfg.dim, task_classes=cfg.task_classes, precision=precision)
    fill_state_dict_(net, seed=0)
    return cfg, net.to(DEV).train()


def run_backward(precision):
    gold = golden_grads()
    cfg, net = build(precision)
    tr = net.trainer()
    x = synthetic_input(tuple(gold["shape"]), seed=0).to(DEV)
    tid = torch.tensor(gold["task_id"]).to(DEV)
    keep = {k: v.to(DEV).contiguous() for k, v in keep_multipliers(cfg, x.shape[0]).items()}
    tr._ensure_packed()
    tr.zero_grad()
    out = torch.empty_like(x)
    with torch.no_grad():
        F = tr.forward_train(x, tr.task_weights(tid), out, keep)
        tr.backward(F, objective_weights(x.shape).to(DEV).contiguous())
    torch.cuda.synchronize()
    return net, tr, out


def grad_errors(net, ref_grads):
    errs = {}
    for n, p in net.named_parameters():
        r = ref_grads.get(n)
        if r is None:
            continue
        g = p.grad.detach().cpu().double()
        errs[n] = float((g - r.double()).norm() / r.double().norm().clamp_min(1e-30))
    return errs


# Per-tensor relative L2 error of the gradient.  fp32 mode = bf16 hi/lo split products (~2^-16 relative): measured
# max 2.1e-4 / median 2.8e-5.  bf16 mode rounds every GEMM operand of a 22-block network to 8 bits of mantissa, forward
# and backward: measured median 2e-2, worst tensors (first latent block) 0.23 — the bound below is per tensor, the
# median bound catches a systematic error.
@pytest.mark.parametrize("precision,out_tol,grad_tol,median_tol", [("fp32", 1e-4, 2e-3, 2e-4), ("bf16", 1e-2, 0.35, 5e-2)])
def test_backward_matches_oracle(precision, out_tol, grad_tol, median_tol):
    import statistics
    ref_out, ref_grads = oracle_grads()
    net, tr, out = run_backward(precision)
    assert rel_err(out.cpu(), ref_out) < out_tol
    errs = grad_errors(net, ref_grads)
    assert len(errs) == 617
    bad = sorted(((e, n) for n, e in errs.items() if not e < grad_tol), reverse=True)
    assert not bad, f"{len(bad)} parameter gradients off, worst: {bad[:8]}"
    assert statistics.median(errs.values()) < median_tol
    # the 8 parameters the reference never touches keep no gradient and stay outside the flat range
    dead = [n for n, p in net.named_parameters() if "text_linear" in n or "clip_linear" in n]
    assert len(dead) == 8 and all(n not in tr.g for n in dead)


def test_bf16_gradients_are_no_worse_than_the_reference_autocast():
    """Yardstick for the bf16 training mode: the UNMODIFIED reference under torch.autocast(bfloat16) — what its own
    `precision="16-mixed"` trainer computes (train.py:118) — is itself 2e-2 (median) / 0.27 (worst tensor) away from its fp32
    gradients on this case (tests/golden/nat_b2_32_grads_bf16_yardstick.json, oracle/make_golden_bf16_yardstick.py).  The
    CUDA trainer's bf16 mode must not be worse: distribution-wise (median, 90th percentile, maximum) and per tensor (within
    a factor of the reference's own error for that tensor, with the reference's 90th percentile as the floor)."""
    import json
    import os
    import statistics
    from tests.conftest import GOLDEN
    with open(os.path.join(GOLDEN, "nat_b2_32_grads_bf16_yardstick.json")) as f:
        yard = json.load(f)
    _, ref_grads = oracle_grads()
    net, tr, _ = run_backward("bf16")
    errs = grad_errors(net, ref_grads)
    ours = sorted(errs.values())
    summ = yard["summary"]
    med, p90, mx = statistics.median(ours), ours[int(0.9 * len(ours))], ours[-1]
    print(f"bf16 gradient rel-L2: ours median {med:.3e} p90 {p90:.3e} max {mx:.3e} | reference autocast median "
          f"{summ['median']:.3e} p90 {summ['p90']:.3e} max {summ['max']:.3e}")
    assert med <= 1.25 * summ["median"] and p90 <= 1.25 * summ["p90"] and mx <= 1.25 * summ["max"]
    worse = {n: (e, yard["rel_l2"][n]) for n, e in errs.items() if e > max(4.0 * yard["rel_l2"][n], summ["p90"])}
    assert not worse, sorted(worse.items(), key=lambda kv: -kv[1][0])[:8]


def test_backward_is_additive_and_zero_grad_resets():
    net, tr, _ = run_backward("fp32")
    g1 = tr.flat_g.clone()
    assert float(g1.abs().max()) > 0
    tr.zero_grad()
    assert float(tr.flat_g.abs().max()) == 0.0


def test_train_step_matches_torch_adamw_on_the_oracle():
    """One full step (clamp+L1, train.py:58-61; AdamW, train.py:69) vs autograd + torch.optim.AdamW on the oracle."""
    cfg, net = build("fp32")
    B = 2
    x = synthetic_input((B, 31, 32, 32), seed=3)
    clean = synthetic_input((B, 31, 32, 32), seed=4)
    tid = torch.tensor([[2], [5]])
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.named_parameters()
          if "text_linear" not in k and "clip_linear" not in k}
    # scale the output conv down so that a good share of the restored values lies inside (0,1) and carries gradient
    with torch.no_grad():
        net.output.weight.mul_(0.05)
        sd["output.weight"].mul_(0.05)
    net.invalidate_packed_weights()
    opt = torch.optim.AdamW(list(sd.values()), lr=2e-4)
    out = O.forward(sd, cfg, x, tid, synthetic_clip_prompt(cfg.task_classes))
    loss = torch.nn.functional.l1_loss(out.clamp(0, 1), clean)
    loss.backward()
    opt.step()
    tr = net.trainer(lr=2e-4)
    got = tr.train_step(x.to(DEV), clean.to(DEV), tid.to(DEV), keep=None)
    torch.cuda.synchronize()
    assert abs(float(got) - float(loss)) < 1e-4 * max(1.0, abs(float(loss)))
    inside = float(((out > 0) & (out < 1)).float().mean())
    assert inside > 0.2, inside
    # parameter agreement after the step (update is lr-sized; sign flips only where |g| ~ 0)
    diffs = []
    for n, p in net.named_parameters():
        if n in sd:
            diffs.append(float((p.detach().cpu() - sd[n].detach()).abs().max()))
    assert max(diffs) <= 2.2 * 2e-4, max(diffs)
    frac_close = sum(d < 2e-5 for d in diffs) / len(diffs)
    assert frac_close > 0.5, frac_close


def test_loss_decreases_over_steps():
    cfg, net = build("bf16")
    with torch.no_grad():
        net.output.weight.mul_(0.05)
    B = 4
    clean = synthetic_input((B, 31, 32, 32), seed=5).to(DEV)
    noisy = (clean + 0.2 * torch.randn(clean.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(0)))
    tid = torch.zeros(B, 1, dtype=torch.long, device=DEV)
    tr = net.trainer(lr=2e-4)
    losses = [float(tr.train_step(noisy, clean, tid)) for _ in range(12)]
    assert all(l == l for l in losses)
    assert sum(losses[-3:]) < sum(losses[:3]), losses
    # inference through the same module sees the updated weights
    net.eval()
    with torch.no_grad():
        y = net(noisy, tid)
    assert torch.isfinite(y).all()


def test_remote_sensing_backward_matches_oracle():
    """wide spectral path (dim 96: head dims 48/96, ranks 12/24, 100 bands, unfused MLP shapes)."""
    from tests.helpers import synthetic_state_dict
    cfg = NetConfig.remote_sensing()
    net = MP_HSIR_Net(cfg.in_channel, cfg.out_channel, cfg.dim, task_classes=cfg.task_classes, precision="fp32")
    fill_state_dict_(net, seed=0)
    net = net.to(DEV).train()
    x = synthetic_input((1, 100, 32, 32), seed=2)
    tid = torch.tensor([[3]])
    R = objective_weights(x.shape, seed=5)
    sd = {k: v.clone().requires_grad_(True) for k, v in synthetic_state_dict("remote_sensing").items()}
    out_ref = O.forward(sd, cfg, x, tid, synthetic_clip_prompt(cfg.task_classes))
    (out_ref * R).sum().backward()
    tr = net.trainer()
    tr._ensure_packed()
    tr.zero_grad()
    out = torch.empty(x.shape, device=DEV)
    with torch.no_grad():
        F = tr.forward_train(x.to(DEV), tr.task_weights(tid.to(DEV)), out, None)
        tr.backward(F, R.to(DEV).contiguous())
    torch.cuda.synchronize()
    assert rel_err(out.cpu(), out_ref.detach()) < 1e-4
    errs = grad_errors(net, {k: v.grad for k, v in sd.items()})
    bad = sorted(((e, n) for n, e in errs.items() if not e < 2e-3), reverse=True)
    assert len(errs) == 617 and not bad, f"{len(bad)} off, worst {bad[:8]}"


def test_cuda_graph_step_matches_eager_step():
    """the captured step (pack / fwd+bwd / AdamW graphs) follows the eager step's loss trajectory"""
    B = 2
    clean = synthetic_input((B, 31, 32, 32), seed=6).to(DEV)
    noisy = clean + 0.15 * torch.randn(clean.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    tid = torch.tensor([[0], [3]], device=DEV)
    traj = {}
    for mode in (False, True):
        cfg, net = build("fp32")
        with torch.no_grad():
            net.output.weight.mul_(0.05)
        tr = net.trainer(lr=2e-4)
        traj[mode] = [float(tr.train_step(noisy, clean, tid, keep=None, cuda_graph=mode)) for _ in range(6)]
        if mode:
            assert "_train_graphs" in tr.__dict__ and tr.step_count == 6
            net.eval()
            with torch.no_grad():
                y = net(noisy, tid)          # inference sees the weights the captured optimiser wrote
            assert torch.isfinite(y).all()
    for a, b in zip(traj[False], traj[True]):
        assert abs(a - b) < 2e-4 * max(abs(a), 1e-3), traj
    assert traj[True][-1] < traj[True][0]


def test_checkpoint_resume_continues_training(tmp_path):
    """save (weights + AdamW moments) -> load into a fresh module -> the next step matches the uninterrupted run"""
    from mp_hsir_b200.checkpoint import load_reference_checkpoint, save_reference_checkpoint
    B = 2
    clean = synthetic_input((B, 31, 32, 32), seed=8).to(DEV)
    noisy = clean + 0.1 * torch.randn(clean.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(2))
    tid = torch.tensor([[1], [2]], device=DEV)
    cfg, net = build("fp32")
    with torch.no_grad():
        net.output.weight.mul_(0.05)
    tr = net.trainer(lr=2e-4)
    for _ in range(3):
        tr.train_step(noisy, clean, tid, keep=None)
    path = str(tmp_path / "resume.ckpt")
    save_reference_checkpoint(path, net, epoch=0, global_step=3, trainer=tr)
    want = float(tr.train_step(noisy, clean, tid, keep=None))
    cfg2, net2 = build("fp32")
    tr2 = net2.trainer(lr=1.0)  # lr / step count come from the checkpoint
    rep = load_reference_checkpoint(path, net2, trainer=tr2)
    assert len(rep["loaded"]) == 658 and tr2.step_count == 3 and abs(tr2.lr - 2e-4) < 1e-12
    got = float(tr2.train_step(noisy, clean, tid, keep=None))
    assert abs(got - want) < 1e-5 * max(abs(want), 1e-3), (got, want)
    w1 = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    w2 = torch.cat([p.detach().reshape(-1) for p in net2.parameters()])
    # same weights, moments and step count going in; the step itself is not bit-reproducible (fp32 atomics in the weight-
    # gradient kernels change the summation order per run) and AdamW's m/sqrt(v) amplifies that jitter where |g| ~ 0:
    # bound the worst weight by a quarter of one lr-sized update (measured: 0.5-0.7 %) and the average by rounding noise — lost
    # moments or a wrong step count would move EVERY weight by a sizeable part of lr (mean ~1e-4)
    d = (w1 - w2).abs()
    assert float(d.max()) < 0.25 * 2e-4 and float(d.mean()) < 1e-7, (float(d.max()), float(d.mean()))


def test_train_step_at_a_non_square_resolution():
    """64x96 patches: no bilinear resize at level 1 rows but one along the columns, 12 x 8 windows, batch 3"""
    cfg, net = build("bf16")
    with torch.no_grad():
        net.output.weight.mul_(0.05)
    B = 3
    clean = synthetic_input((B, 31, 64, 96), seed=9).to(DEV)
    noisy = clean + 0.1 * torch.randn(clean.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    tid = torch.tensor([[0], [2], [5]], device=DEV)
    tr = net.trainer(lr=2e-4)
    losses = [float(tr.train_step(noisy, clean, tid, cuda_graph=(i >= 2))) for i in range(8)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses


def test_backward_batch_of_one_with_1d_task_id():
    """B = 1 and an eval-style 1-D task id (Text_Prompt's one-hot branch, net/MP_HSIR.py:525) through the backward"""
    from tests.helpers import synthetic_state_dict
    cfg, net = build("fp32")
    x = synthetic_input((1, 31, 32, 64), seed=12)
    tid = torch.tensor([4])
    R = objective_weights(x.shape, seed=13)
    sd = {k: v.clone().requires_grad_(True) for k, v in synthetic_state_dict("natural").items()}
    out_ref = O.forward(sd, cfg, x, tid, synthetic_clip_prompt(cfg.task_classes))
    (out_ref * R).sum().backward()
    tr = net.trainer()
    tr._ensure_packed()
    tr.zero_grad()
    out = torch.empty(x.shape, device=DEV)
    with torch.no_grad():
        F = tr.forward_train(x.to(DEV), tr.task_weights(tid.to(DEV)), out, None)
        tr.backward(F, R.to(DEV).contiguous())
    torch.cuda.synchronize()
    assert rel_err(out.cpu(), out_ref.detach()) < 1e-4
    errs = grad_errors(net, {k: v.grad for k, v in sd.items()})
    bad = sorted(((e, n) for n, e in errs.items() if not e < 2e-3), reverse=True)
    assert len(errs) == 617 and not bad, f"{len(bad)} off, worst {bad[:8]}"


def test_captured_step_owns_every_buffer_it_writes():
    """regression: the restored batch / its gradient are written by every replay of the forward+backward graph, so they
    must stay allocated for the graph's lifetime — tensors a user allocates after the capture must never be clobbered"""
    B = 2
    clean = synthetic_input((B, 31, 32, 32), seed=14).to(DEV)
    noisy = clean + 0.1 * torch.randn(clean.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(4))
    tid = torch.tensor([[0], [1]], device=DEV)
    cfg, net = build("bf16")
    tr = net.trainer(lr=2e-4)
    for _ in range(2):                                   # eager warm-up step, then capture + first replay
        tr.train_step(noisy, clean, tid, keep=None, cuda_graph=True)
    torch.cuda.synchronize()
    sentinels = [torch.full_like(noisy, 7.0) for _ in range(8)]   # same size class as the graph's output buffers
    for _ in range(3):
        tr.train_step(noisy, clean, tid, keep=None, cuda_graph=True)
    torch.cuda.synchronize()
    assert all(bool((s == 7.0).all()) for s in sentinels)


def _train_py_step(net, opt, x, clean, tid):
    """PromptIRModel.training_step + the optimiser step Lightning wraps around it (train.py:50-69), verbatim."""
    restored = net(x, tid)                                   # train.py:58
    restored = torch.clamp(restored, 0, 1)                   # train.py:59
    loss = torch.nn.functional.l1_loss(restored, clean)      # train.py:61 (nn.L1Loss)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.detach()


def test_module_drops_into_train_py_autograd_loop():
    """The drop-in module under autograd — net(x, t) in train mode, loss.backward(), torch.optim.AdamW(net.parameters()) —
    against net.trainer().train_step (hand-written backward + fused AdamW) on the same batches with the same DropPath
    draws: losses to 1e-6, first-step gradients to 5e-4 relative L2 per tensor (the two paths run the SAME kernels; the
    weight-gradient atomics make the summation order vary), parameters after two steps equal except where |g| ~ 0
    (AdamW's first updates are lr * sign(g))."""
    steps, lr = 2, 2e-4
    xs = [synthetic_input((2, 31, 32, 32), seed=40 + i).to(DEV) for i in range(steps)]
    cs = [synthetic_input((2, 31, 32, 32), seed=50 + i).to(DEV) for i in range(steps)]
    tid = torch.tensor([[1], [4]]).to(DEV)

    cfg, net_a = build("fp32")
    with torch.no_grad():
        net_a.output.weight.mul_(0.05)
    opt = torch.optim.AdamW(net_a.parameters(), lr=lr)       # train.py:69 — created BEFORE the first forward, like Lightning
    cfg, net_b = build("fp32")
    with torch.no_grad():
        net_b.output.weight.mul_(0.05)
    tr = net_b.trainer(lr=lr)
    gen_b = torch.Generator(device=DEV).manual_seed(123)

    losses_a, losses_b, grad_errs = [], [], {}
    for i, (x, c) in enumerate(zip(xs, cs)):
        if i == 0:
            net_a.trainer().drop_path_generator = torch.Generator(device=DEV).manual_seed(123)
        losses_a.append(float(_train_py_step(net_a, opt, x, c, tid)))
        keep = tr.drop_path_scales(x.shape[0], gen_b)
        if i == 0:
            # gradients of the first step, before either optimiser has moved anything
            tr.zero_grad()
            _, l = tr.loss_and_grad(x, c, tid, keep=keep)
            for n, p in net_a.named_parameters():
                if p.grad is not None:
                    g = tr.g[n].view(p.shape)
                    grad_errs[n] = float((p.grad - g).norm() / g.norm().clamp_min(1e-30))
            losses_b.append(float(l))
            tr.optimizer_step()
        else:
            losses_b.append(float(tr.train_step(x, c, tid, keep=keep)))
    torch.cuda.synchronize()
    # step 1 starts from identical parameters: the losses agree to rounding; later steps inherit AdamW's sign flips on
    # near-zero gradients (the weight-gradient atomics make their summation order, hence their last bits, vary per run)
    assert abs(losses_a[0] - losses_b[0]) <= 1e-6, (losses_a, losses_b)
    assert all(abs(a - b) <= 1e-4 for a, b in zip(losses_a, losses_b)), (losses_a, losses_b)
    assert len(grad_errs) == 617 and max(grad_errs.values()) < 5e-4, sorted(grad_errs.items(), key=lambda kv: -kv[1])[:5]
    dead = [n for n, p in net_a.named_parameters() if p.grad is None]
    assert sorted(dead) == sorted(n for n, _ in net_a.named_parameters() if "text_linear" in n or "clip_linear" in n)
    pa, pb = dict(net_a.named_parameters()), dict(net_b.named_parameters())
    diffs = [float((pa[n] - pb[n]).abs().max()) for n in pa]
    assert max(diffs) <= 2.2 * lr * steps, max(diffs)
    import statistics
    assert statistics.median(diffs) < 2e-5, sorted(diffs)[-20:]   # a tensor's max |difference|: one flipped element costs 2 lr = 4e-4
    y = net_a(xs[0], tid)
    assert y.requires_grad and y.grad_fn is not None


def test_second_forward_invalidates_the_first_graph():
    cfg, net = build("fp32")
    x = synthetic_input((1, 31, 32, 32), seed=1).to(DEV)
    tid = torch.tensor([0]).to(DEV)
    y1 = net(x, tid)
    y2 = net(x, tid)
    with pytest.raises(RuntimeError, match="LATEST forward"):
        y1.sum().backward()
    y2.sum().backward()
    assert net.output.weight.grad is not None and torch.isfinite(net.output.weight.grad).all()


def test_captured_steps_survive_a_larger_inference_forward():
    """a forward at a larger shape re-allocates named workspace buffers the captured training step points into;
    the captured graphs must be dropped and re-captured, not replayed on freed memory."""
    cfg, net = build("fp32")
    with torch.no_grad():
        net.output.weight.mul_(0.05)
    tr = net.trainer(lr=2e-4)
    x = synthetic_input((2, 31, 32, 32), seed=60).to(DEV)
    c = synthetic_input((2, 31, 32, 32), seed=61).to(DEV)
    tid = torch.tensor([[0], [2]]).to(DEV)
    for _ in range(3):                        # eager warm-up, capture, replay
        tr.train_step(x, c, tid, keep=None, cuda_graph=True)
    gen0 = tr.ws.generation
    net.eval()
    with torch.no_grad():
        big = tr.forward(synthetic_input((1, 31, 64, 96), seed=62).to(DEV), torch.tensor([1]).to(DEV))
    net.train()
    assert tr.ws.generation != gen0 and torch.isfinite(big).all()
    l_graph = float(tr.train_step(x, c, tid, keep=None, cuda_graph=True))      # must not replay the stale graph
    # same state, eager, on a twin
    cfg, net2 = build("fp32")
    with torch.no_grad():
        net2.output.weight.mul_(0.05)
    tr2 = net2.trainer(lr=2e-4)
    for _ in range(3):
        tr2.train_step(x, c, tid, keep=None)
    l_eager = float(tr2.train_step(x, c, tid, keep=None))
    assert abs(l_graph - l_eager) <= 1e-5, (l_graph, l_eager)
