"""CPU oracle for the MP-HSIR hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch functional restatement (plain PyTorch on CPU tensors, channels-last,
no nn.Module) of the forward pass of ``MP_HSIR_Net`` in the reference
``net/MP_HSIR.py``.  Every function cites the reference lines it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this file; the product path (``mp_hsir_b200``) never does.

Pinning: the reference ships no tests or golden vectors for this path,
so the pin is generated — ``oracle/make_golden.py`` imports the UNMODIFIED reference
module from /root/reference (behind ``timm``/``clip`` stubs), fills it with the
name-seeded weights of ``mp_hsir_b200.synth`` and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against those
vectors.  The CLIP text encoder (third-party ``clip`` package, not vendored, weights are
a network download) is replaced by a synthetic [T,512] tensor on both sides: that one
input is *parity unpinned* and documented as such in DESIGN.md.

Data layout used here: activations are ``[B, H, W, C]`` (token-major, the layout
PGSSTB uses internally, net/MP_HSIR.py:665-668); weights are taken verbatim from the
reference ``state_dict`` (a ``dict[str, Tensor]``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from mp_hsir_b200.config import LN_EPS, PROMPT_LEN, SHIFT, WINDOW, NetConfig, Stage

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """LayerNorm over the last (channel) axis, biased variance, eps 1e-5.

    nn.LayerNorm(C) on tokens (net/MP_HSIR.py:618-619) and WithBias_LayerNorm
    (net/MP_HSIR.py:354-357) are the same arithmetic.
    """
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def gelu(x: torch.Tensor) -> torch.Tensor:
    """exact erf GELU (nn.GELU() default, net/MP_HSIR.py:67; F.gelu, :263,:389)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def conv1x1(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """bias-free 1x1 conv on channels-last data; w is [Cout, Cin, 1, 1]."""
    return x @ w.reshape(w.shape[0], w.shape[1]).t()


def conv3x3(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """bias-free dense 3x3 conv, zero pad 1; x [B,H,W,Cin], w [Cout,Cin,3,3]."""
    B, H, W, _ = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    out = x.new_zeros(B, H, W, w.shape[0])
    for dy in range(3):
        for dx in range(3):
            out = out + xp[:, dy:dy + H, dx:dx + W, :] @ w[:, :, dy, dx].t()
    return out


def dwconv3x3(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """bias-free depthwise 3x3 conv, zero pad 1; x [B,H,W,C], w [C,1,3,3]."""
    B, H, W, _ = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    out = torch.zeros_like(x)
    for dy in range(3):
        for dx in range(3):
            out = out + xp[:, dy:dy + H, dx:dx + W, :] * w[:, 0, dy, dx]
    return out


def pixel_unshuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelUnshuffle(2) on channels-last: out[..., c*4+2i+j] = in[2h+i, 2w+j, c] (net/MP_HSIR.py:437)."""
    B, H, W, C = x.shape
    x = x.view(B, H // 2, 2, W // 2, 2, C)            # b h i w j c
    return x.permute(0, 1, 3, 5, 2, 4).reshape(B, H // 2, W // 2, C * 4)


def pixel_shuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelShuffle(2) on channels-last (net/MP_HSIR.py:447)."""
    B, H, W, C4 = x.shape
    C = C4 // 4
    x = x.view(B, H, W, C, 2, 2)                      # b h w c i j
    return x.permute(0, 1, 4, 2, 5, 3).reshape(B, H * 2, W * 2, C)


def to_windows(x: torch.Tensor, shift: int) -> torch.Tensor:
    """roll(-s,-s) + window_partition -> [B*nW, 64, C] (net/MP_HSIR.py:21-30, 671-678)."""
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    B, H, W, C = x.shape
    x = x.view(B, H // WINDOW, WINDOW, W // WINDOW, WINDOW, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, WINDOW * WINDOW, C)


def from_windows(w: torch.Tensor, shift: int, B: int, H: int, W: int) -> torch.Tensor:
    """window_reverse + roll(+s,+s) -> [B,H,W,C] (net/MP_HSIR.py:33-44, 689-696)."""
    C = w.shape[-1]
    x = w.view(B, H // WINDOW, W // WINDOW, WINDOW, WINDOW, C)
    x = x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    return x


def shift_mask(H: int, W: int, dtype=torch.float32) -> torch.Tensor:
    """Swin shifted-window mask [nW,64,64] in {0,-100}, closed form of calculate_mask
    (net/MP_HSIR.py:639-660): region label 3*rh+rw in *shifted* coordinates."""
    ys = torch.arange(H)
    xs = torch.arange(W)
    rh = (ys >= H - WINDOW).long() + (ys >= H - SHIFT).long()
    rw = (xs >= W - WINDOW).long() + (xs >= W - SHIFT).long()
    lab = (3 * rh[:, None] + rw[None, :]).view(1, H, W, 1).to(dtype)
    lw = lab.view(1, H // WINDOW, WINDOW, W // WINDOW, WINDOW, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64)
    diff = lw[:, None, :] - lw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def relative_position_bias(table: torch.Tensor) -> torch.Tensor:
    """[heads,64,64] bias from the [225,heads] table; index (yp-yq+7)*15+(xp-xq+7)
    (net/MP_HSIR.py:172-181, 200-202)."""
    t = torch.arange(64)
    y, x = t // 8, t % 8
    idx = (y[:, None] - y[None, :] + 7) * 15 + (x[:, None] - x[None, :] + 7)
    return table[idx.reshape(-1)].view(64, 64, -1).permute(2, 0, 1).contiguous()


# --------------------------------------------------------------------------------------
# PGSSTB pieces
# --------------------------------------------------------------------------------------


def window_attention_core(xw: torch.Tensor, sd: SD, p: str, heads: int,
                          mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Spatial_Attention up to (not including) proj: [B_,64,C] -> [B_,64,C]
    (net/MP_HSIR.py:193-215)."""
    B_, N, C = xw.shape
    hd = C // heads
    qkv = xw @ sd[p + "qkv.weight"].t() + sd[p + "qkv.bias"]
    qkv = qkv.view(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1) + relative_position_bias(sd[p + "relative_position_bias_table"])[None]
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask[None, :, None]).view(B_, heads, N, N)
    attn = attn.softmax(-1)
    return (attn @ v).transpose(1, 2).reshape(B_, N, C)


def spatial_attention(xw, sd, p, heads, mask):
    """full Spatial_Attention.forward (net/MP_HSIR.py:193-218)."""
    o = window_attention_core(xw, sd, p, heads, mask)
    return o @ sd[p + "proj.weight"].t() + sd[p + "proj.bias"]


def local_spectral_gate(sa_mean: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """Per-window channel gate g[B_,C] of PG_Spectral_Attention; its output is sa*g
    (net/MP_HSIR.py:132-155).  sa_mean is the token mean of the window [B_,C]."""
    r = sd[p + "linear_down.weight"].shape[0]
    pw = F.softmax(sa_mean @ sd[p + "linear_prompt.weight"].t(), dim=-1)          # [B_,128]
    down = sa_mean @ sd[p + "linear_down.weight"].t()                              # [B_,r]
    sp = pw @ sd[p + "prompt_param"].view(PROMPT_LEN, r)                           # [B_,r]
    q = sp @ sd[p + "q.weight"].t()
    kv = down @ sd[p + "kv.weight"].t()
    k, v = kv[:, :r], kv[:, r:]
    a = (q[:, :, None] * k[:, None, :]) * (r ** -0.5)                              # [B_,r,r] outer product
    a = a.softmax(-1)
    o = (a * v[:, None, :]).sum(-1)                                                # [B_,r]
    o = o @ sd[p + "proj.weight"].t() + sd[p + "proj.bias"]
    return o @ sd[p + "linear_up.weight"].t()                                      # [B_,C]


def transposed_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                         temperature: torch.Tensor, heads: int) -> torch.Tensor:
    """Channel ("transposed") attention over all tokens. q,k,v [B,T,C] -> [B,T,C]
    (net/MP_HSIR.py:101-112, 239-247, 412-424)."""
    B, T, C = q.shape
    c = C // heads

    def split(t):
        return t.view(B, T, heads, c).permute(0, 2, 3, 1)                          # b head c T

    q, k, v = split(q), split(k), split(v)
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    attn = (q @ k.transpose(-2, -1)) * temperature.view(1, heads, 1, 1)
    attn = attn.softmax(-1)
    out = attn @ v                                                                 # b head c T
    return out.permute(0, 3, 1, 2).reshape(B, T, C)


def global_spectral_attention(x: torch.Tensor, sd: SD, p: str, heads: int) -> torch.Tensor:
    """Spectral_Attention / MDTA Attention: [B,H,W,C] -> [B,H,W,C]
    (net/MP_HSIR.py:96-114, identical math at :406-427)."""
    B, H, W, C = x.shape
    qkv = dwconv3x3(conv1x1(x, sd[p + "qkv.weight"]), sd[p + "qkv_dwconv.weight"])
    q, k, v = qkv.reshape(B, H * W, 3 * C).split(C, dim=-1)
    out = transposed_attention(q, k, v, sd[p + "temperature"], heads)
    return conv1x1(out, sd[p + "project_out.weight"]).view(B, H, W, C)


def gated_mlp(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """GatedMlp: value = first half of fc1, gate = second half (net/MP_HSIR.py:76-82)."""
    h = x @ sd[p + "fc1.weight"].t() + sd[p + "fc1.bias"]
    hid = h.shape[-1] // 2
    h = h[..., :hid] * gelu(h[..., hid:])
    return h @ sd[p + "fc2.weight"].t() + sd[p + "fc2.bias"]


def gdfn(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """FeedForward / FFN (GDFN): gelu(first half) * second half (net/MP_HSIR.py:386-391, 260-265)."""
    h = dwconv3x3(conv1x1(x, sd[p + "project_in.weight"]), sd[p + "dwconv.weight"])
    hid = h.shape[-1] // 2
    return conv1x1(gelu(h[..., :hid]) * h[..., hid:], sd[p + "project_out.weight"])


def pgsstb(x: torch.Tensor, sd: SD, p: str, heads: int, shift: int,
           keep: Optional[torch.Tensor] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """One PGSSTB block on channels-last data [B,H,W,C] (net/MP_HSIR.py:662-723).

    ``keep`` (optional [2,B] tensor) is the DropPath multiplier (mask/keep_prob) of the two
    residual branches for train-mode parity (net/MP_HSIR.py:718-719); None = eval.
    ``taps`` (optional dict) receives intermediates for per-kernel tests.
    """
    B, H, W, C = x.shape
    shortcut = x
    xn = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    xw = to_windows(xn, shift)
    # the reference builds a mask whenever (H,W) differs from the construction resolution,
    # all-zero for un-shifted blocks (net/MP_HSIR.py:680-683); adding zeros is a no-op.
    mask = shift_mask(H, W, x.dtype) if shift else None
    core = window_attention_core(xw, sd, p + "attn.", heads, mask)
    sa_w = core @ sd[p + "attn.proj.weight"].t() + sd[p + "attn.proj.bias"]        # [B_,64,C]
    gate = local_spectral_gate(sa_w.mean(dim=1), sd, p + "local_spectral_attn.")   # [B_,C]
    x1 = from_windows(sa_w * gate[:, None, :], shift, B, H, W)
    sa = from_windows(sa_w, shift, B, H, W)
    x2 = global_spectral_attention(sa, sd, p + "gobal_spectral_attn.", heads)
    y = x1 + x2
    if keep is not None:
        y = y * keep[0].view(B, 1, 1, 1)
    x = shortcut + y
    m = gated_mlp(layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), sd, p + "mlp.")
    if keep is not None:
        m = m * keep[1].view(B, 1, 1, 1)
    out = x + m
    if taps is not None:
        taps.update(xn=xn, core=from_windows(core, shift, B, H, W), core_win_mean=core.mean(dim=1),
                    sa=sa, gate=gate, x1=x1, x2=x2, mid=x, out=out)
    return out


def base_block(x: torch.Tensor, sd: SD, st: Stage, keep=None) -> torch.Tensor:
    """BaseBlock: x + blocks(x); shift 0 on even, 4 on odd blocks (net/MP_HSIR.py:746-761)."""
    y = x
    for i in range(st.depth):
        k = None if keep is None else keep.get((st.name, i))  # blocks with rate 0 hold nn.Identity (:620)
        y = pgsstb(y, sd, f"{st.name}.blocks.{i}.", st.heads, SHIFT if i % 2 else 0, k)
    return y + x


# --------------------------------------------------------------------------------------
# prompts
# --------------------------------------------------------------------------------------


def text_prompt(task_id: torch.Tensor, clip_prompt: torch.Tensor, task_classes: int):
    """Text_Prompt.forward (net/MP_HSIR.py:517-532): returns (clip[B,512], weights[B,T]).

    1-D task_id -> one-hot; 2-D task_id [B,k] -> mean of the k one-hots.  The class mean
    divides by T, so the result is sum_t w[b,t]*clip[t]/T.
    """
    if task_id.dim() > 1:
        w = F.one_hot(task_id.long(), task_classes).to(clip_prompt.dtype).mean(dim=1)
    else:
        w = F.one_hot(task_id.long(), task_classes).to(clip_prompt.dtype)
    return (w @ clip_prompt) / task_classes, w


def nearest_index(n_out: int, n_in: int) -> torch.Tensor:
    """F.interpolate(mode='nearest') source index: floor(dst * n_in / n_out)."""
    return torch.clamp((torch.arange(n_out, dtype=torch.float32) * (n_in / n_out)).floor().long(), max=n_in - 1)


def bilinear_resize(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """F.interpolate(mode='bilinear', align_corners=False) on channels-last data
    (net/MP_HSIR.py:580): src = max((dst+0.5)*in/out - 0.5, 0)."""
    B, h, w, C = x.shape
    if (h, w) == (H, W):
        return x

    def axis(n_out, n_in):
        s = (torch.arange(n_out, dtype=torch.float32) + 0.5) * (n_in / n_out) - 0.5
        s = s.clamp_min(0.0)
        i0 = s.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        lam = (s - i0.to(torch.float32)).to(x.dtype)
        return i0, i1, lam

    y0, y1, ly = axis(H, h)
    x0, x1, lx = axis(W, w)
    rows = x[:, y0] * (1 - ly).view(1, H, 1, 1) + x[:, y1] * ly.view(1, H, 1, 1)
    return rows[:, :, x0] * (1 - lx).view(1, 1, W, 1) + rows[:, :, x1] * lx.view(1, 1, W, 1)


def tvsp_query(clip_b: torch.Tensor, weights: torch.Tensor, learnable: torch.Tensor, ps: int) -> torch.Tensor:
    """The text-side query map of TVSP [B,ps,ps,D] (net/MP_HSIR.py:575-577).

    tp[b,d] = mean_t(w[b,t]*learnable[t,d]); ``tp[B,D,1,1] * clip[B,512]`` broadcasts to
    [B,D,B,512] and is nearest-resized to [ps,ps], so
    Q[b,i,j,d] = tp[b,d] * clip[floor(i*B/ps), floor(j*512/ps)]  — rows index *other samples*.
    """
    B = clip_b.shape[0]
    T = weights.shape[1]
    tp = (weights.to(learnable.dtype) @ learnable.view(T, -1)) / T                 # [B,D]
    src = clip_b[nearest_index(ps, B)][:, nearest_index(ps, clip_b.shape[1])]      # [ps,ps]
    return tp[:, None, None, :] * src[None, :, :, None]


def cross_transformer(xq: torch.Tensor, xkv: torch.Tensor, sd: SD, p: str, heads: int = 2) -> torch.Tensor:
    """CrossTransformer with cross_residual=True (net/MP_HSIR.py:280-287, 234-249)."""
    B, H, W, D = xq.shape
    nq = layer_norm(xq, sd[p + "norm11.body.weight"], sd[p + "norm11.body.bias"])
    nkv = layer_norm(xkv, sd[p + "norm12.body.weight"], sd[p + "norm12.body.bias"])
    q = dwconv3x3(conv1x1(nq, sd[p + "attn.q.weight"]), sd[p + "attn.q_dwconv.weight"])
    kv = dwconv3x3(conv1x1(nkv, sd[p + "attn.kv.weight"]), sd[p + "attn.kv_dwconv.weight"])
    k, v = kv.reshape(B, H * W, 2 * D).split(D, dim=-1)
    o = transposed_attention(q.reshape(B, H * W, D), k, v, sd[p + "attn.temperature"], heads)
    x = xq + conv1x1(o, sd[p + "attn.project_out.weight"]).view(B, H, W, D)
    return x + gdfn(layer_norm(x, sd[p + "norm2.body.weight"], sd[p + "norm2.body.bias"]), sd, p + "ffn.")


def tvsp(H: int, W: int, clip_b: torch.Tensor, weights: torch.Tensor, sd: SD, p: str, ps: int) -> torch.Tensor:
    """TVSP.forward: only the *shape* of the feature map is used (net/MP_HSIR.py:572-583)."""
    B = clip_b.shape[0]
    T = weights.shape[1]
    q = tvsp_query(clip_b, weights, sd[p + "text_prompt_learnable"].view(T, -1), ps)
    vis = sd[p + "visual_prompt"].permute(0, 2, 3, 1).expand(B, -1, -1, -1)
    pr = cross_transformer(q, vis, sd, p + "cross_transformer.")
    return conv3x3(bilinear_resize(pr, H, W), sd[p + "conv_last.weight"])


def transformer_block(x: torch.Tensor, sd: SD, p: str, heads: int) -> torch.Tensor:
    """Restormer TransformerBlock: x+MDTA(LN(x)); x+GDFN(LN(x)) (net/MP_HSIR.py:475-479)."""
    x = x + global_spectral_attention(
        layer_norm(x, sd[p + "norm1.body.weight"], sd[p + "norm1.body.bias"]), sd, p + "attn.", heads)
    return x + gdfn(layer_norm(x, sd[p + "norm2.body.weight"], sd[p + "norm2.body.bias"]), sd, p + "ffn.")


def prompt_fusion(x: torch.Tensor, prompt: torch.Tensor, sd: SD, p: str, heads: int) -> torch.Tensor:
    """PromptFusion: cat -> TransformerBlock -> conv1x1 (net/MP_HSIR.py:594-599)."""
    y = transformer_block(torch.cat([x, prompt], dim=-1), sd, p + "transformer.", heads)
    return conv1x1(y, sd[p + "conv.weight"])


# --------------------------------------------------------------------------------------
# whole network
# --------------------------------------------------------------------------------------


def forward(sd: SD, cfg: NetConfig, inp: torch.Tensor, task_id: torch.Tensor,
            clip_prompt: torch.Tensor, keep=None, taps: Optional[dict] = None) -> torch.Tensor:
    """MP_HSIR_Net.forward (net/MP_HSIR.py:810-844). inp is NCHW like the reference; so is the result."""
    st = {s.name: s for s in cfg.stages()}
    clip_b, w = text_prompt(task_id, clip_prompt.to(inp.dtype), cfg.task_classes)
    x = inp.permute(0, 2, 3, 1)
    B, H, W, _ = x.shape

    x1 = conv3x3(x, sd["patch_embed.proj.weight"])
    e1 = base_block(x1, sd, st["encoder_level1"], keep)
    x2 = pixel_unshuffle2(conv3x3(e1, sd["down1_2.body.0.weight"]))
    e2 = base_block(x2, sd, st["encoder_level2"], keep)
    x3 = pixel_unshuffle2(conv3x3(e2, sd["down2_3.body.0.weight"]))
    lat = base_block(x3, sd, st["latent"], keep)

    d2 = pixel_shuffle2(conv3x3(lat, sd["up3_2.body.0.weight"]))
    p2 = tvsp(H // 2, W // 2, clip_b, w, sd, "prompt2.", 32)
    f2 = prompt_fusion(e2, p2, sd, "fusion2.", 8)
    d2 = conv1x1(torch.cat([d2, f2], dim=-1), sd["reduce_chan_level2.weight"])
    d2 = base_block(d2, sd, st["decoder_level2"], keep)

    d1 = pixel_shuffle2(conv3x3(d2, sd["up2_1.body.0.weight"]))
    p1 = tvsp(H, W, clip_b, w, sd, "prompt1.", 64)
    f1 = prompt_fusion(e1, p1, sd, "fusion1.", 4)
    d1 = torch.cat([d1, f1], dim=-1)
    d1 = base_block(d1, sd, st["decoder_level1"], keep)
    d1 = base_block(d1, sd, st["refinement"], keep)
    out = conv3x3(d1, sd["output.weight"]) + x
    if taps is not None:
        taps.update(x1=x1, e1=e1, x2=x2, e2=e2, x3=x3, lat=lat, p2=p2, f2=f2, p1=p1, f1=f1, d1=d1)
    return out.permute(0, 3, 1, 2).contiguous()


def psnr(restored: torch.Tensor, clean: torch.Tensor) -> float:
    """Per-band PSNR (data_range 1) on clip(.,0,1), mean over bands then batch
    (utils/val_utils.py:49-69 of the reference; skimage's peak_signal_noise_ratio formula)."""
    r = restored.detach().double().clamp(0, 1)
    c = clean.detach().double().clamp(0, 1)
    mse = ((r - c) ** 2).mean(dim=(-2, -1))
    return float((10.0 * torch.log10(1.0 / mse)).mean(dim=1).mean())
