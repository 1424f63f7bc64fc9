"""GPU: the tcgen05 GEMM engine (gemm_tc.cu) through the C ABI against fp64 torch.

bf16x3 (hi/lo split operands) is held to 5e-5 per op (measured product error ~2^-16; whole network
1.1e-5 in the CPU emulation); bf16 to 1.5e-2 per op.
"""
import pytest
import torch

from mp_hsir_b200 import engine as E
from mp_hsir_b200 import lib
from mp_hsir_b200.lib import View, Weight
from oracle import mp_hsir_oracle as O
from tests.conftest import rel_err
from tests.test_ops_gpu import DEV, V, dev, out_mat, rnd, tokens

pytestmark = pytest.mark.gpu
TOL = {lib.PREC_BF16X3: 5e-5, lib.PREC_BF16: 1.5e-2}
PRECS = [lib.PREC_BF16X3, lib.PREC_BF16]


def weight(w_nk: torch.Tensor) -> Weight:
    """logical [N,K] fp32 -> Weight with only the tensor-core image."""
    n, k = w_nk.shape
    return Weight(None, lib.pack_bimg(dev(w_nk), n, k), n, k)


def mm64(a, w):
    return (a.double() @ w.double().t()).float()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (256, 64, 192), (4096, 128, 384), (300, 96, 288), (1000, 176, 64),
                                   (640, 352, 128), (512, 688, 256), (20000, 128, 704), (384, 1024, 384), (128, 256, 1376)])
def test_gemm_tc_bias(prec, M, K, N):
    a, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    y = out_mat(M, N)
    lib.gemm(V(dev(a)), weight(w), V(y), N, bias=dev(b), precision=prec)
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), mm64(a, w) + b) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("K", [64, 128, 256, 384])
def test_gemm_tc_layernorm(prec, K):
    M, N = 1000, 3 * K
    a, w, b = rnd(M, K, seed=4) * 3 + 1, rnd(N, K, seed=5, scale=K ** -0.5), rnd(N, seed=6)
    g, be = 1 + 0.1 * rnd(K, seed=7), 0.1 * rnd(K, seed=8)
    y = out_mat(M, N)
    lib.gemm(V(dev(a)), weight(w), V(y), N, ln=(dev(g), dev(be)), bias=dev(b), precision=prec)
    ref = mm64(O.layer_norm(a.double(), g.double(), be.double()).float(), w) + b
    assert rel_err(y.cpu(), ref) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("C,hid", [(64, 170), (128, 340), (96, 255)])
def test_gemm_tc_gated_mlp(prec, C, hid):
    M = 1000
    hp = E._ceil(hid, 16)
    sd = {"fc1.weight": rnd(2 * hid, C, seed=1, scale=C ** -0.5), "fc1.bias": 0.1 * rnd(2 * hid, seed=2),
          "fc2.weight": rnd(C, hid, seed=3, scale=hid ** -0.5), "fc2.bias": 0.1 * rnd(C, seed=4)}
    g, be = 1 + 0.1 * rnd(C, seed=7), 0.1 * rnd(C, seed=8)
    x = rnd(M, C, seed=9)
    w1, b1 = E.pack_glu_fc1(sd["fc1.weight"], sd["fc1.bias"], hid, hp)
    w2 = E.pack_linear_t(sd["fc2.weight"], k_pad=hp)
    W1 = weight(w1[:C, :2 * hp].t().contiguous())
    W2 = weight(w2[:hp, :C].t().contiguous())
    xd = dev(x)
    h, y = out_mat(M, hp), out_mat(M, C)
    lib.gemm(V(xd), W1, V(h), 2 * hp, ln=(dev(g), dev(be)), bias=dev(b1), epi=lib.EPI_GLU, precision=prec)
    lib.gemm(V(h), W2, V(y), C, bias=dev(sd["fc2.bias"]), epi=lib.EPI_RESIDUAL, res1=V(xd), precision=prec)
    ref = (x.double() + O.gated_mlp(O.layer_norm(x.double(), g.double(), be.double()),
                                    {k: v.double() for k, v in sd.items()}, "")).float()
    assert torch.isfinite(h).all()
    assert rel_err(y.cpu(), ref) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
def test_gemm_tc_per_sample_weights_shared_operand_and_spectral_epilogue(prec):
    B, H, W, C = 3, 8, 24, 64      # hw = 192: not a multiple of the 128-row tile
    hw = H * W
    a = rnd(hw, C, seed=1)          # shared by all samples
    wb = rnd(B, C, C, seed=2, scale=C ** -0.5)  # per-sample logical [N,K]
    r1, sa = rnd(B * hw, C, seed=3), rnd(B * hw, C, seed=4)
    gate = rnd(B * hw // 64, C, seed=5)
    img = lib.pack_bimg(dev(wb), C, C)
    y = out_mat(B * hw, C)
    lib.gemm(V(dev(a)), Weight(None, img, C, C), V(y), C, epi=lib.EPI_SPECTRAL, res1=V(dev(r1)), gsrc=V(dev(sa)),
             gate=dev(gate), H=H, W=W, shift=4, rows_per_batch=hw, a_row_mod=hw, M=B * hw, precision=prec)
    x2 = torch.cat([mm64(a, wb[b]) for b in range(B)])
    g_img = O.from_windows(gate[:, None, :].expand(-1, 64, -1), 4, B, H, W).reshape(B * hw, C)
    ref = r1 + sa * g_img + x2
    assert rel_err(y.cpu(), ref) < TOL[prec]


@pytest.mark.parametrize("prec", [lib.PREC_FP32_SIMT] + PRECS)
@pytest.mark.parametrize("C,shift", [(64, 4), (96, 0), (256, 4)])
def test_gemm_proj_epilogue_split_outputs(prec, C, shift):
    """MPHSIR_EPI_PROJ: left n_split columns -> res1 + s_b * (acc+bias) * gate[window]; the rest -> Y2."""
    B, H, W = 2, 16, 24
    hw = H * W
    a = rnd(B * hw, C, seed=1)
    w = rnd(4 * C, C, seed=2, scale=C ** -0.5)
    bias, r1 = rnd(4 * C, seed=3), rnd(B * hw, C, seed=4)
    gate = rnd(B * hw // 64, C, seed=5)
    scale = torch.tensor([1.0, 0.5])
    y, y2 = out_mat(B * hw, C), out_mat(B * hw, 3 * C)
    wobj = Weight(E.pack_linear_t(dev(w)), None if prec == lib.PREC_FP32_SIMT else lib.pack_bimg(dev(w), 4 * C, C), 4 * C, C)
    lib.gemm(V(dev(a)), wobj, V(y), 4 * C, bias=dev(bias), epi=lib.EPI_PROJ, res1=V(dev(r1)), gate=dev(gate), Y2=V(y2),
             n_split=C, H=H, W=W, shift=shift, rows_per_batch=hw, row_scale=dev(scale), precision=prec)
    full = mm64(a, w) + bias
    g_img = O.from_windows(gate[:, None, :].expand(-1, 64, -1), shift, B, H, W).reshape(B * hw, C)
    s_row = scale.repeat_interleave(hw)[:, None]
    tol = 2e-6 if prec == lib.PREC_FP32_SIMT else TOL[prec]
    assert rel_err(y.cpu(), r1 + s_row * full[:, :C] * g_img) < tol
    assert rel_err(y2.cpu(), full[:, C:]) < tol


@pytest.mark.parametrize("prec", PRECS)
def test_conv3x3_tc_all_output_modes(prec):
    B, H, W, C = 2, 16, 24, 64
    x = rnd(B, H, W, C, seed=1)
    xd = dev(tokens(x))
    # tokens
    w = rnd(96, C, 3, 3, seed=2, scale=(9 * C) ** -0.5)
    y = out_mat(B * H * W, 96)
    lib.conv3x3(V(xd), weight(E.pack_conv3x3(w)[:, :96].t().contiguous()), y.data_ptr(), 96, B, H, W, C, 96, precision=prec)
    assert rel_err(y.cpu(), tokens(O.conv3x3(x.double(), w.double()).float())) < TOL[prec]
    # unshuffle
    wd = rnd(C // 2, C, 3, 3, seed=3, scale=(9 * C) ** -0.5)
    y = out_mat(B * H * W // 4, 2 * C)
    lib.conv3x3(V(xd), weight(E.pack_conv3x3(wd)[:, :C // 2].t().contiguous()), y.data_ptr(), 2 * C, B, H, W, C, C // 2,
                lib.CONV_UNSHUFFLE, precision=prec)
    assert rel_err(y.cpu(), tokens(O.pixel_unshuffle2(O.conv3x3(x.double(), wd.double()).float()))) < TOL[prec]
    # shuffle into the left half of a concat buffer
    wu = rnd(2 * C, C, 3, 3, seed=4, scale=(9 * C) ** -0.5)
    cat = torch.zeros(B * H * W * 4, C, device=DEV)
    lib.conv3x3(V(xd), weight(E.pack_conv3x3(wu, shuffle=True)[:, :2 * C].t().contiguous()), cat.data_ptr(), C, B, H, W, C,
                2 * C, lib.CONV_SHUFFLE, precision=prec)
    assert rel_err(cat[:, : C // 2].cpu(), tokens(O.pixel_shuffle2(O.conv3x3(x.double(), wu.double()).float()))) < TOL[prec]
    assert (cat[:, C // 2:] == 0).all()
    # 31-band NCHW in / NCHW + residual out
    img = rnd(1, 31, 32, 32, seed=5)
    tok = out_mat(32 * 32, 32)
    lib.nchw_to_tokens(dev(img), V(tok))
    w1 = rnd(64, 31, 3, 3, seed=6, scale=(9 * 31) ** -0.5)
    f = out_mat(32 * 32, 64)
    lib.conv3x3(V(tok), weight(E.pack_conv3x3(w1, cin_pad=32)[:, :64].t().contiguous()), f.data_ptr(), 64, 1, 32, 32, 32, 64,
                precision=prec)
    ref = O.conv3x3(img.double().permute(0, 2, 3, 1), w1.double())
    assert rel_err(f.cpu(), tokens(ref.float())) < TOL[prec]
    w2 = rnd(31, 64, 3, 3, seed=7, scale=(9 * 64) ** -0.5)
    out = torch.full((1, 31, 32, 32), float("nan"), device=DEV)
    lib.conv3x3(V(f), weight(E.pack_conv3x3(w2)[:, :31].t().contiguous()), out.data_ptr(), 0, 1, 32, 32, 64, 31,
                lib.CONV_NCHW_RES, R=dev(img), precision=prec)
    ref2 = (O.conv3x3(ref, w2.double()).permute(0, 3, 1, 2) + img.double()).float()
    assert rel_err(out.cpu(), ref2) < TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("C,hid,M", [(64, 170, 1000), (128, 340, 128), (128, 340, 5000), (64, 170, 40000), (128, 340, 150000)])
def test_fused_mlp_matches_gated_mlp(prec, C, hid, M):
    """one-kernel LN -> fc1 -> value*gelu(gate) -> fc2 -> x + s*(.) + res2 (net/MP_HSIR.py:719, :76-82)."""
    hp = E._ceil(hid, 16)
    assert lib.mlp_supported(C, hp)
    sd = {"fc1.weight": rnd(2 * hid, C, seed=1, scale=C ** -0.5), "fc1.bias": 0.1 * rnd(2 * hid, seed=2),
          "fc2.weight": rnd(C, hid, seed=3, scale=hid ** -0.5), "fc2.bias": 0.1 * rnd(C, seed=4)}
    g, be = 1 + 0.1 * rnd(C, seed=7), 0.1 * rnd(C, seed=8)
    x, r2 = rnd(M, C, seed=9), rnd(M, C, seed=10)
    w1, b1 = E.pack_glu_fc1(sd["fc1.weight"], sd["fc1.bias"], hid, hp)
    w2 = E.pack_linear_t(sd["fc2.weight"], k_pad=hp)
    W1 = weight(w1[:C, :2 * hp].t().contiguous())
    W2 = weight(w2[:hp, :C].t().contiguous())
    y = out_mat(M, C)
    lib.mlp(V(dev(x)), (dev(g), dev(be)), W1, dev(b1), W2, dev(sd["fc2.bias"]), V(y), hp, prec, res2=V(dev(r2)))
    torch.cuda.synchronize()
    ref = (x.double() + O.gated_mlp(O.layer_norm(x.double(), g.double(), be.double()),
                                    {k: v.double() for k, v in sd.items()}, "") + r2.double()).float()
    assert rel_err(y.cpu(), ref) < TOL[prec]
