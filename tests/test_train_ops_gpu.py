"""GPU: every backward / training kernel of the C ABI against PyTorch-CPU autograd of the oracle's op."""
import pytest
import torch
import torch.nn.functional as F

from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View
from oracle import mp_hsir_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dev(t):
    return t.to(DEV).contiguous()


def V(t):
    return View.of(t)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(*shape, generator=g)


@pytest.fixture(params=[1, 0], ids=["wgrad_tcgen05", "wgrad_mma_sync"], autouse=True)
def wgrad_engine(request):
    """every test of this file runs with both weight-gradient kernels"""
    lib.load().mphsir_debug_wgrad_tc(request.param)
    yield request.param
    lib.load().mphsir_debug_wgrad_tc(-1)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [(lib.PREC_BF16X3, 2e-5), (lib.PREC_BF16, 8e-3)])
@pytest.mark.parametrize("M,O_,I", [(1000, 192, 64), (4096, 352, 128), (130, 8, 16), (777, 136, 24)])
def test_wgrad_plain(prec, tol, M, O_, I):
    dY, X = rnd(M, O_, seed=1), rnd(M, I, seed=2)
    dW = torch.zeros(O_ * I, device=DEV)
    lib.wgrad(V(dev(dY)), V(dev(X)), dW, prec)
    assert rel(dW.view(O_, I), dY.double().t() @ X.double()) < tol
    # accumulates
    lib.wgrad(V(dev(dY)), V(dev(X)), dW, prec)
    assert rel(dW.view(O_, I), 2 * (dY.double().t() @ X.double())) < tol


def test_wgrad_maps_and_padding():
    M, hid, hp, C = 600, 170, 176, 64
    X = rnd(M, C, seed=3)
    # interleaved fc1 layout: packed col 2j = value_j (param row j), 2j+1 = gate_j (param row hid + j)
    dH = rnd(M, 2 * hp, seed=4)
    dW = torch.zeros(2 * hid * C, device=DEV)
    lib.wgrad(V(dev(dH)), V(dev(X)), dW, lib.PREC_BF16X3, map_mode=lib.MAP_INTERLEAVE, map_a=hid)
    full = dH.double().t() @ X.double()
    ref = torch.cat([full[0:2 * hid:2], full[1:2 * hid:2]], 0)
    assert rel(dW.view(2 * hid, C), ref) < 2e-5
    # padded halves (GDFN): [0,hid) and [hp, hp+hid)
    dW2 = torch.zeros(2 * hid * C, device=DEV)
    lib.wgrad(V(dev(dH)), V(dev(X)), dW2, lib.PREC_BF16X3, map_mode=lib.MAP_HALVES, map_a=hid, map_b=hp)
    ref2 = torch.cat([full[:hid], full[hp:hp + hid]], 0)
    assert rel(dW2.view(2 * hid, C), ref2) < 2e-5
    # padded input columns (fc2: X = hidden [M, hp], param [C, hid])
    Hd, dYc = rnd(M, hp, seed=5), rnd(M, C, seed=6)
    dW3 = torch.zeros(C * hid, device=DEV)
    lib.wgrad(V(dev(dYc)), V(dev(Hd)), dW3, lib.PREC_BF16X3, i_valid=hid)
    assert rel(dW3.view(C, hid), (dYc.double().t() @ Hd.double())[:, :hid]) < 2e-5
    # bias gradient with the same maps: stand-alone column sums, and riding on the weight-gradient launch
    db = torch.zeros(2 * hid, device=DEV)
    lib.colsum(V(dev(dH)), db, lib.MAP_INTERLEAVE, hid)
    s = dH.double().sum(0)
    assert rel(db, torch.cat([s[0:2 * hid:2], s[1:2 * hid:2]])) < 1e-5
    db2, dW4 = torch.zeros(2 * hid, device=DEV), torch.zeros(2 * hid * C, device=DEV)
    lib.wgrad(V(dev(dH)), V(dev(X)), dW4, lib.PREC_BF16X3, map_mode=lib.MAP_INTERLEAVE, map_a=hid, dbias=db2)
    assert rel(db2, torch.cat([s[0:2 * hid:2], s[1:2 * hid:2]])) < 1e-5 and rel(dW4.view(2 * hid, C), ref) < 2e-5
    # wide layer (several 128-row tiles, two input-channel tiles): every column is summed exactly once
    M2, O2, I2 = 3000, 384, 192
    dY2, X2 = rnd(M2, O2, seed=60), rnd(M2, I2, seed=61)
    db3, dW5 = torch.zeros(O2, device=DEV), torch.zeros(O2 * I2, device=DEV)
    lib.wgrad(V(dev(dY2)), V(dev(X2)), dW5, lib.PREC_BF16X3, dbias=db3)
    assert rel(db3, dY2.double().sum(0)) < 1e-5 and rel(dW5.view(O2, I2), dY2.double().t() @ X2.double()) < 2e-5


def test_wgrad_multi_problem_launch():
    M = 700
    shapes = [(128, 64), (8, 64), (128, 8), (8, 8), (16, 8), (64, 8), (136, 24)]
    probs, refs, outs = [], [], []
    db = torch.zeros(8, device=DEV)
    for j, (O_, I) in enumerate(shapes):
        dY, X = rnd(M, O_, seed=70 + j), rnd(M, I, seed=90 + j)
        out = torch.zeros(O_ * I, device=DEV)
        kw = {"dbias": db} if j == 3 else {}
        probs.append((V(dev(dY)), V(dev(X)), out, kw))
        refs.append(dY.double().t() @ X.double())
        outs.append(out)
        if j == 3:
            bias_ref = dY.double().sum(0)
    lib.wgrad_multi(probs, lib.PREC_BF16X3)
    for (O_, I), out, ref in zip(shapes, outs, refs):
        assert rel(out.view(O_, I), ref) < 2e-5, (O_, I)
    assert rel(db, bias_ref) < 1e-5


def test_wgrad_batched_and_shared():
    B, HW, C = 3, 320, 64
    dU, v = rnd(B * HW, C, seed=7), rnd(B * HW, C, seed=8)
    P = torch.zeros(B * C * C, device=DEV)
    lib.wgrad(V(dev(dU)), V(dev(v)), P, lib.PREC_BF16X3, so=C, rows_per_batch=HW, dw_batch_stride=C * C)
    ref = torch.einsum("bno,bnj->boj", dU.view(B, HW, C).double(), v.view(B, HW, C).double())
    assert rel(P.view(B, C, C), ref) < 2e-5
    vs = rnd(HW, C, seed=9)
    P.zero_()
    lib.wgrad(V(dev(dU)), V(dev(vs)), P, lib.PREC_BF16X3, so=C, rows_per_batch=HW, dw_batch_stride=C * C, x_row_mod=HW)
    ref = torch.einsum("bno,nj->boj", dU.view(B, HW, C).double(), vs.double())
    assert rel(P.view(B, C, C), ref) < 2e-5


@pytest.mark.parametrize("Cin,Cout,cin_valid", [(32, 64, 31), (64, 32, 64), (128, 48, 128)])
def test_wgrad_conv3x3(Cin, Cout, cin_valid):
    B, H, W = 2, 16, 24
    x = rnd(B, H, W, Cin, seed=10)
    x[..., cin_valid:] = 0
    w = rnd(Cout, cin_valid, 3, 3, seed=11).requires_grad_(True)
    dY = rnd(B, H, W, Cout, seed=12)
    y = O.conv3x3(x[..., :cin_valid], w)
    (y * dY).sum().backward()
    dW = torch.zeros(Cout * cin_valid * 9, device=DEV)
    lib.wgrad(V(dev(dY.view(-1, Cout))), V(dev(x.view(-1, Cin))), dW, lib.PREC_BF16X3, taps=9, H=H, W=W,
              so=9 * cin_valid, si=9, st=1, map_a=Cout, i_valid=cin_valid)
    assert rel(dW.view(Cout, cin_valid, 3, 3), w.grad) < 2e-5


def test_layernorm_fwd_bwd():
    M, C = 515, 192
    x = rnd(M, C, seed=13).requires_grad_(True)
    g_, b_ = (1 + 0.1 * rnd(C, seed=14)).requires_grad_(True), (0.1 * rnd(C, seed=15)).requires_grad_(True)
    G, add = rnd(M, C, seed=16), rnd(M, C, seed=17)
    y = O.layer_norm(x, g_, b_)
    (y * G).sum().backward()
    xd = dev(x.detach())
    Y, stats = torch.empty(M, C, device=DEV), torch.empty(2 * M, device=DEV)
    lib.layernorm_fwd(V(xd), (dev(g_.detach()), dev(b_.detach())), V(Y), stats)
    assert rel(Y, y.detach()) < 1e-5
    dX, dg, db = torch.empty(M, C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    lib.layernorm_bwd(V(xd), stats, dev(g_.detach()), V(dev(G)), V(dev(add)), V(dX), dg, db)
    assert rel(dX, x.grad + add) < 1e-5
    assert rel(dg, g_.grad) < 1e-5 and rel(db, b_.grad) < 1e-5


def test_glu_and_gdfn_gate_bwd():
    M, hid, hp = 300, 170, 176
    val, gate = rnd(M, hp, seed=18).requires_grad_(True), rnd(M, hp, seed=19).requires_grad_(True)
    dh = rnd(M, hp, seed=20)
    hidden = val * O.gelu(gate)
    (hidden * dh).sum().backward()
    H = torch.stack([val.detach(), gate.detach()], -1).reshape(M, 2 * hp)      # interleaved
    Hd, Dd = dev(H), dev(dh)
    lib.glu_bwd(V(Hd), V(Dd), hp)
    assert rel(Dd, hidden.detach()) < 1e-5
    got = Hd.cpu().view(M, hp, 2)
    assert rel(got[..., 0], val.grad) < 1e-5 and rel(got[..., 1], gate.grad) < 1e-5
    # GDFN: gelu(a) * b, halves
    a, b = rnd(M, hp, seed=21).requires_grad_(True), rnd(M, hp, seed=22).requires_grad_(True)
    y = O.gelu(a) * b
    (y * dh).sum().backward()
    T = dev(torch.cat([a.detach(), b.detach()], 1))
    Y = torch.empty(M, hp, device=DEV)
    lib.gdfn_gate_fwd(V(T), V(Y), hp)
    assert rel(Y, y.detach()) < 1e-5
    dT = torch.empty(M, 2 * hp, device=DEV)
    lib.gdfn_gate_bwd(V(T), V(dev(dh)), V(dT), hp)
    assert rel(dT[:, :hp], a.grad) < 1e-5 and rel(dT[:, hp:], b.grad) < 1e-5


def test_axpby_batch_sum():
    B, rows, C = 3, 50, 64
    x, y = rnd(B * rows, C, seed=23), rnd(B * rows, C, seed=24)
    s = torch.tensor([0.0, 1.25, 2.0])
    yd = dev(y)
    lib.axpby(V(dev(x)), V(yd), 0.5, 2.0, row_scale=dev(s), rows_per_batch=rows)
    ref = 0.5 * s.repeat_interleave(rows)[:, None] * x + 2.0 * y
    assert rel(yd, ref) < 1e-6
    xs = rnd(rows, C, seed=25)
    lib.axpby(V(dev(xs)), V(yd), x_row_mod=rows, M=B * rows)
    assert rel(yd, xs.repeat(B, 1)) < 1e-7
    out = torch.empty(rows, C, device=DEV)
    lib.batch_sum(V(dev(x)), V(out), B)
    assert rel(out, x.view(B, rows, C).sum(0)) < 1e-6


@pytest.mark.parametrize("prec,tol", [(lib.PREC_FP32_SIMT, 1e-5), (lib.PREC_BF16X3, 5e-5), (lib.PREC_BF16, 2e-2)])
@pytest.mark.parametrize("C,heads,shift,H,W", [(64, 2, 0, 16, 16), (64, 2, 4, 16, 24), (128, 4, 4, 16, 16), (96, 2, 4, 16, 16),
                                               (192, 2, 4, 8, 16)])
def test_window_attn_bwd(C, heads, shift, H, W, prec, tol):
    B = 2
    hd = C // heads
    qkv = rnd(B, H, W, 3 * C, seed=26, scale=0.7).requires_grad_(True)
    table = (0.5 * rnd(225, heads, seed=27)).requires_grad_(True)
    dO = rnd(B, H, W, C, seed=28)
    xw = O.to_windows(qkv, shift)                                   # [B_,64,3C]
    B_ = xw.shape[0]
    t = xw.view(B_, 64, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = t[0] * hd ** -0.5, t[1], t[2]
    attn = q @ k.transpose(-2, -1) + O.relative_position_bias(table)[None]
    if shift:
        mask = O.shift_mask(H, W)
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, 64, 64) + mask[None, :, None]).view(B_, heads, 64, 64)
    core = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B_, 64, C)
    out = O.from_windows(core, shift, B, H, W)
    (out * dO).sum().backward()
    qd = dev(qkv.detach().view(-1, 3 * C))
    rpb = dev(O.relative_position_bias(table.detach()))
    # forward kernel output must agree too (same qkv)
    dq = torch.empty(B * H * W, 3 * C, device=DEV)
    groups = lib.window_attn_bwd_groups(B, H, W, heads)
    partial = torch.empty(groups * heads * 4096, device=DEV)
    lib.window_attn_bwd(V(qd), rpb, V(dev(dO.view(-1, C))), V(dq), partial, groups, B, H, W, C, heads, shift, precision=prec)
    assert rel(dq, qkv.grad.view(-1, 3 * C)) < tol
    dbias = torch.zeros(heads * 4096, device=DEV)
    lib.colsum(View(partial.data_ptr(), heads * 4096, groups, heads * 4096, partial), dbias)
    dtab = torch.zeros(225 * heads, device=DEV)
    lib.rpb_table_bwd(dbias, dtab, heads)
    assert rel(dtab.view(225, heads), table.grad) < tol


@pytest.mark.parametrize("C,r,shift", [(64, 8, 0), (128, 16, 4), (192, 12, 4)])
def test_local_gate_backward_chain(C, r, shift):
    """window_reduce + LL GEMM + local_gate_bwd + record wgrads + gate_apply_bwd == autograd of sa * gate[win]."""
    B, H, W = 2, 16, 16
    N, B_ = B * H * W, B * H * W // 64
    p = "l."
    sd = {
        p + "linear_prompt.weight": rnd(128, C, seed=30, scale=C ** -0.5), p + "linear_down.weight": rnd(r, C, seed=31, scale=C ** -0.5),
        p + "prompt_param": torch.rand(1, 1, 128, r, generator=torch.Generator().manual_seed(32)),
        p + "q.weight": rnd(r, r, seed=33, scale=0.5), p + "kv.weight": rnd(2 * r, r, seed=34, scale=0.5),
        p + "proj.weight": rnd(r, r, seed=35, scale=0.5), p + "proj.bias": rnd(r, seed=36, scale=0.1),
        p + "linear_up.weight": rnd(C, r, seed=37, scale=0.5),
    }
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    sa = rnd(B, H, W, C, seed=38).requires_grad_(True)
    dU = rnd(B, H, W, C, seed=39)
    sa_w = O.to_windows(sa, shift)
    gate = O.local_spectral_gate(sa_w.mean(1), sd, p)
    x1 = O.from_windows(sa_w * gate[:, None, :], shift, B, H, W)
    (x1 * dU).sum().backward()

    prec = lib.PREC_BF16X3
    sad, dUd = dev(sa.detach().view(N, C)), dev(dU.view(N, C))
    dg, msa = torch.empty(B_ * C, device=DEV), torch.empty(B_ * C, device=DEV)
    lib.window_reduce(V(dUd), V(sad), dg, B, H, W, C, shift, 1.0)
    lib.window_reduce(V(sad), None, msa, B, H, W, C, shift, 1.0 / 64)
    assert rel(msa.view(B_, C), sa_w.detach().mean(1)) < 1e-6
    cat = torch.cat([sd[p + "linear_prompt.weight"], sd[p + "linear_down.weight"]], 0).detach()      # [128+r, C]
    LL = dev(sa_w.detach().mean(1) @ cat.t())                                                        # checked op: GEMM engine elsewhere
    raw = {"param": dev(sd[p + "prompt_param"].detach().view(128, r)), "q": dev(sd[p + "q.weight"].detach()),
           "kv": dev(sd[p + "kv.weight"].detach()), "proj": dev(sd[p + "proj.weight"].detach()),
           "proj_bias": dev(sd[p + "proj.bias"].detach()), "up": dev(sd[p + "linear_up.weight"].detach())}
    ldr = lib.local_gate_bwd_record_ld(r)
    rec = torch.zeros(B_, ldr, device=DEV)
    lib.local_gate_bwd(V(LL), dg, raw, V(rec), B_, C, r)
    recv = V(rec)
    dmean = (rec[:, :128 + r].cpu().double() @ cat.double())                                          # [B_, C]
    msa_v, dg_v = View(msa.data_ptr(), C, B_, C, msa), View(dg.data_ptr(), C, B_, C, dg)
    o_w, o_dsp, o_dq, o_sp, o_dkv, o_low, o_du, o_o, o_u = (128 + r, 256 + r, 256 + 2 * r, 256 + 3 * r, 256 + 4 * r,
                                                             256 + 6 * r, 256 + 7 * r, 256 + 8 * r, 256 + 9 * r)

    def wg(dy, x, shape):
        out = torch.zeros(shape[0] * shape[1], device=DEV)
        lib.wgrad(dy, x, out, prec)
        return out.view(shape)

    tol = 2e-4
    assert rel(wg(recv.cols_slice(0, 128), msa_v, (128, C)), sd[p + "linear_prompt.weight"].grad) < tol
    assert rel(wg(recv.cols_slice(128, 128 + r), msa_v, (r, C)), sd[p + "linear_down.weight"].grad) < tol
    assert rel(wg(recv.cols_slice(o_w, o_w + 128), recv.cols_slice(o_dsp, o_dsp + r), (128, r)), sd[p + "prompt_param"].grad.view(128, r)) < tol
    assert rel(wg(recv.cols_slice(o_dq, o_dq + r), recv.cols_slice(o_sp, o_sp + r), (r, r)), sd[p + "q.weight"].grad) < tol
    assert rel(wg(recv.cols_slice(o_dkv, o_dkv + 2 * r), recv.cols_slice(o_low, o_low + r), (2 * r, r)), sd[p + "kv.weight"].grad) < tol
    assert rel(wg(recv.cols_slice(o_du, o_du + r), recv.cols_slice(o_o, o_o + r), (r, r)), sd[p + "proj.weight"].grad) < tol
    assert rel(wg(dg_v, recv.cols_slice(o_u, o_u + r), (C, r)), sd[p + "linear_up.weight"].grad) < tol
    db = torch.zeros(r, device=DEV)
    lib.colsum(recv.cols_slice(o_du, o_du + r), db)
    assert rel(db, sd[p + "proj.bias"].grad) < tol
    dsa = torch.empty(N, C, device=DEV)
    lib.gate_apply_bwd(V(dUd), dev(gate.detach()), dev(dmean.float()), V(dsa), B, H, W, C, shift)
    assert rel(dsa, sa.grad.view(N, C)) < tol


def test_dwconv_wgrad_and_flipped_dgrad():
    B, H, W, C = 2, 12, 20, 96
    x = rnd(B, H, W, C, seed=40).requires_grad_(True)
    w = rnd(C, 1, 3, 3, seed=41).requires_grad_(True)
    dY = rnd(B, H, W, C, seed=42)
    (O.dwconv3x3(x, w) * dY).sum().backward()
    dW = torch.zeros(C * 9, device=DEV)
    xd, dYd = dev(x.detach().view(-1, C)), dev(dY.view(-1, C))
    lib.dwconv3x3_wgrad(V(xd), V(dYd), dW, B, H, W, C)
    assert rel(dW.view(C, 1, 3, 3), w.grad) < 1e-5
    w9 = dev(w.detach().reshape(C, 9).t())
    dX = torch.empty(B * H * W, C, device=DEV)
    lib.dwconv3x3(V(dYd), w9.flip(0).contiguous(), V(dX), B, H, W, C)
    assert rel(dX, x.grad.view(-1, C)) < 1e-5
    # halves map (GDFN dwconv [2*hid] packed at [0,hid) / [hp,hp+hid))
    hid, hp = 40, 48
    dW2 = torch.zeros(2 * hid * 9, device=DEV)
    lib.dwconv3x3_wgrad(V(xd), V(dYd), dW2, B, H, W, C, lib.MAP_HALVES, hid, hp)
    assert rel(dW2.view(2 * hid, 9), torch.cat([w.grad.view(C, 9)[:hid], w.grad.view(C, 9)[hp:hp + hid]], 0)) < 1e-5


def test_pixel_shuffles_and_layouts():
    B, H, W, C = 2, 8, 12, 20
    x = rnd(B, H, W, C, seed=43)
    out = torch.empty(B * H * W // 4, 4 * C, device=DEV)
    lib.pixel_unshuffle(V(dev(x.view(-1, C))), V(out), B, H, W, C)
    assert torch.equal(out.cpu(), O.pixel_unshuffle2(x).reshape(-1, 4 * C))
    y = rnd(B, H, W, 4 * C, seed=44)
    out2 = torch.empty(B * H * W * 4, C, device=DEV)
    lib.pixel_shuffle(V(dev(y.view(-1, 4 * C))), V(out2), B, H, W, C)
    assert torch.equal(out2.cpu(), O.pixel_shuffle2(y).reshape(-1, C))
    nchw = torch.empty(B, C, H * W, device=DEV)
    lib.tokens_to_nchw(V(dev(x.view(-1, C))), nchw, B, C, H * W)
    assert torch.equal(nchw.cpu().view(B, C, H, W), x.permute(0, 3, 1, 2))


def test_bilinear_and_tvsp_query_bwd():
    B, h, w, H, W, C = 2, 8, 8, 20, 12, 16
    x = rnd(B, h, w, C, seed=45).requires_grad_(True)
    dY = rnd(B, H, W, C, seed=46)
    (O.bilinear_resize(x, H, W) * dY).sum().backward()
    dX = torch.zeros(B * h * w, C, device=DEV)
    lib.bilinear_bwd(V(dev(dY.view(-1, C))), V(dX), B, h, w, H, W, C)
    assert rel(dX, x.grad.view(-1, C)) < 1e-5
    T, D, ps = 6, 64, 16
    clip_b, wts = rnd(B, 512, seed=47), torch.tensor([[0.5, 0, 0.5, 0, 0, 0], [0, 0, 0, 1.0, 0, 0]])
    learn = rnd(T, D, seed=48).requires_grad_(True)
    dQ = rnd(B, ps, ps, D, seed=49)
    (O.tvsp_query(clip_b, wts, learn, ps) * dQ).sum().backward()
    dL = torch.zeros(T * D, device=DEV)
    lib.tvsp_query_bwd(V(dev(dQ.view(-1, D))), dev(clip_b), dev(wts), dL, B, T, D, ps)
    assert rel(dL.view(T, D), learn.grad) < 1e-5


def test_l1_clamp_loss_and_adamw():
    out = (rnd(2, 31, 16, 16, seed=50) * 0.6 + 0.5).requires_grad_(True)
    clean = torch.rand(2, 31, 16, 16, generator=torch.Generator().manual_seed(51))
    loss = F.l1_loss(torch.clamp(out, 0, 1), clean)
    loss.backward()
    dO, lb = torch.empty(out.shape, device=DEV), torch.zeros(1, device=DEV)
    lib.l1_clamp_loss(dev(out.detach()), dev(clean), dO, lb)
    assert abs(float(lb) - float(loss)) < 1e-6 and rel(dO, out.grad) < 1e-6
    n = 1003
    p0, g = rnd(n, seed=52), rnd(n, seed=53, scale=0.01)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=2e-4)
    p, m, v = dev(p0), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        pt.grad = g * step
        opt.step()
        lib.adamw_step(p, dev(g * step * 4.0), m, v, 2e-4, 0.9, 0.999, 1e-8, 1e-2, step, grad_scale=0.25)
    assert rel(p, pt.detach()) < 1e-6
