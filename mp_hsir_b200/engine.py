"""Forward executor: packed weights + workspace + the static kernel sequence of MP_HSIR_Net.forward.

Host logic only (Python): every arithmetic step of the hot path is a libmphsir.so launch
(mp_hsir_b200/lib.py).  torch is used to allocate device memory, to re-lay-out *weights* once
(transpose / pad / interleave — data movement, cached until a parameter changes) and to build the
tiny [B,T] one-hot task weights.

Activation layout: token-major fp32 rows (see include/mphsir.h).  Concatenations of the reference
(net/MP_HSIR.py:595, :827, :835) never materialise: producers write straight into column slices of
the consumer's buffer.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import lib
from .config import PROMPT_LEN, SHIFT, Stage
from .lib import View


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _ldb(n: int) -> int:
    return 64 if n <= 64 else _ceil(n, 128)


# ------------------------------------------------------------------------------------------------
# weight packing (pure data movement; runs once per parameter version)
# ------------------------------------------------------------------------------------------------


def pack_linear_t(w: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """[out,in] (nn.Linear / squeezed 1x1 conv) -> Bt [Kp, ldb] "in x out", zero padded."""
    w = w.reshape(w.shape[0], -1)
    n, k = w.shape
    kp = _ceil(k if k_pad is None else k_pad, 16)
    out = w.new_zeros(kp, _ldb(n))
    out[:k, :n] = w.t()
    return out.contiguous()


def pack_glu_fc1(w: torch.Tensor, b: torch.Tensor, hid: int, hid_pad: int):
    """GatedMlp.fc1 [2*hid, C]: value rows [0,hid), gate rows [hid,2hid) (net/MP_HSIR.py:77) ->
    interleaved columns (2j, 2j+1) = (value_j, gate_j), padded to 2*hid_pad columns."""
    c = w.shape[1]
    n = 2 * hid_pad
    wt = w.new_zeros(_ceil(c, 16), _ldb(n))
    bias = w.new_zeros(_ldb(n))
    wt[:c, 0:2 * hid:2] = w[:hid].t()
    wt[:c, 1:2 * hid:2] = w[hid:].t()
    bias[0:2 * hid:2] = b[:hid]
    bias[1:2 * hid:2] = b[hid:]
    return wt.contiguous(), bias.contiguous()


def pack_gdfn(project_in: torch.Tensor, dw: torch.Tensor, project_out: torch.Tensor, hid: int, hid_pad: int):
    """GDFN (net/MP_HSIR.py:380-391): halves of project_in / dwconv placed at [0,hid) and
    [hid_pad, hid_pad+hid) so the gate kernel sees 16-byte aligned halves."""
    d = project_in.shape[1]
    pin = project_in.reshape(2 * hid, d)
    n = 2 * hid_pad
    wt = pin.new_zeros(_ceil(d, 16), _ldb(n))
    wt[:d, :hid] = pin[:hid].t()
    wt[:d, hid_pad:hid_pad + hid] = pin[hid:].t()
    w9 = pin.new_zeros(9, n)
    dwf = dw.reshape(2 * hid, 9)
    w9[:, :hid] = dwf[:hid].t()
    w9[:, hid_pad:hid_pad + hid] = dwf[hid:].t()
    pout = pack_linear_t(project_out.reshape(d, hid), k_pad=hid_pad)
    return wt.contiguous(), w9.contiguous(), pout


def pack_conv3x3(w: torch.Tensor, cin_pad: Optional[int] = None, shuffle: bool = False) -> torch.Tensor:
    """[Cout,Cin,3,3] -> Wt [9*Cin_p, ldb], row = tap*Cin_p + c, tap = 3*ky + kx.  With
    ``shuffle`` the output columns are permuted to q*Cn+cn (q = 2i+j) so that the PixelShuffle(2)
    store of MPHSIR_CONV_SHUFFLE is a 128-bit write (reference order is cn*4+q, net/MP_HSIR.py:447)."""
    cout, cin = w.shape[:2]
    cp = _ceil(cin if cin_pad is None else cin_pad, 16)
    if shuffle:
        cn = cout // 4
        w = w.view(cn, 4, cin, 3, 3).permute(1, 0, 2, 3, 4).reshape(cout, cin, 3, 3)
    out = w.new_zeros(9, cp, _ldb(cout))
    out[:, :cin, :cout] = w.permute(2, 3, 1, 0).reshape(9, cin, cout)
    return out.reshape(9 * cp, -1).contiguous()


def pack_dw(w: torch.Tensor) -> torch.Tensor:
    """depthwise [C,1,3,3] -> [9, C]."""
    return w.reshape(w.shape[0], 9).t().contiguous()


def rel_pos_bias(table: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """pre-gather [225,heads] -> [heads,64,64] (net/MP_HSIR.py:200-202)."""
    return table[index.reshape(-1)].view(64, 64, -1).permute(2, 0, 1).contiguous()


# ------------------------------------------------------------------------------------------------


PRECISIONS = {"fp32": lib.PREC_BF16X3, "fp32_exact": lib.PREC_FP32_SIMT, "bf16": lib.PREC_BF16}


class Workspace:
    """Named, grow-only fp32 device buffers; a forward at a fixed shape allocates nothing after the
    first call (CUDA-graph friendly)."""

    def __init__(self, device, zero_new: bool = False):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        self.generation = 0  # bumped whenever a buffer is (re)allocated: captured graphs hold raw pointers
        # sharded scenes compute on row sub-ranges of their buffers and copy halos in from neighbours; rows nobody has
        # written yet must at least be finite
        self._new = torch.zeros if zero_new else torch.empty

    def mat(self, name: str, rows: int, cols: int) -> View:
        n = rows * cols
        t = self.bufs.get(name)
        if t is None or t.numel() < n:
            t = self._new(max(n, 1), device=self.device, dtype=torch.float32)
            self.bufs[name] = t
            self.generation += 1
        return View(t.data_ptr(), cols, rows, cols, t)

    def flat(self, name: str, n: int) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < n:
            t = self._new(max(n, 1), device=self.device, dtype=torch.float32)
            self.bufs[name] = t
            self.generation += 1
        return t

    def raw(self, name: str, nbytes: int) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 128), device=self.device, dtype=torch.uint8)
            self.bufs[name] = t
            self.generation += 1
        return t

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class Engine:
    def __init__(self, net):
        self.net = net
        self.cfg = net.cfg
        p0 = next(net.parameters())
        if not p0.is_cuda:
            raise RuntimeError("MP_HSIR_Net parameters must live on a CUDA device before forward()")
        self.device = p0.device
        lib.load()
        with torch.cuda.device(self.device):
            self.sm_count = lib.device_check(self.device.index or 0)
        self.ws = Workspace(self.device)
        self.prec = PRECISIONS[net.precision]
        self.fuse_mlp = True
        self.fold_weights = True   # fp64 weight folds (proj into the gate / spectral-qkv matrices); the trainer turns it off
        self.fold_proj = True
        self.split_gate = True
        self.fused_gate = True   # one-launch local gate (lib.local_gate2); False: GEMM + tail kernel (split_gate)
        self.packed: Optional[dict] = None
        self._versions = None
        self._graphs: Dict[tuple, tuple] = {}
        # TVSP.forward (net/MP_HSIR.py:572-583) depends on the task ids and the SHAPE of the input only, never on its values:
        # at inference its result is kept in the prompt columns of the fusion buffers and re-used while (task ids, B, H, W),
        # the packed weights and the workspace allocation stay the same
        self.cache_prompts = True
        self._tvsp_valid: Dict[str, tuple] = {}
        self._pack_serial = 0
        # trainer only: streams over which a captured re-pack spreads its per-module re-layouts (see _pack_lane)
        self._pack_streams = None
        self._pack_main = None
        # row band of a scene sharded over GPUs (mp_hsir_b200/sharded.py sets it per resolution level): the image the block
        # kernels see is the band plus an 8-row halo on either side
        self.band = None

    # -- weights -------------------------------------------------------------------------------
    def invalidate(self):
        self.packed = None
        self._graphs.clear()
        self._tvsp_valid.clear()

    def _param_versions(self):
        return tuple(p._version for p in self.net.parameters())

    def _ensure_packed(self):
        v = self._param_versions()
        if self.packed is None or v != self._versions:
            # the tensor-core image packs are queued and issued as a handful of multi-matrix launches
            queue_here = lib.PACK_QUEUE is None
            if queue_here:
                lib.PACK_QUEUE = []
            try:
                self.packed = self._pack()
            finally:
                if queue_here:
                    lib.flush_packs()
            self._versions = v
            self._graphs.clear()
            self._tvsp_valid.clear()
            self._pack_serial += 1

    def _pack_lane(self, i: int) -> None:
        """Trainer: the weights change every step and their re-layout (~25 tiny launches per block, ~600 per step) is
        replayed from a CUDA graph.  While that graph is captured the modules are dealt round-robin to a few forked streams
        (lane -1 = back to the capturing stream), so the replay runs the independent re-layouts side by side instead of
        as one 1.2 ms chain of 2 us kernels (13 lanes: 0.41 ms).  No-op outside that capture."""
        if self._pack_streams:
            torch.cuda.set_stream(self._pack_main if i < 0 else self._pack_streams[i % len(self._pack_streams)])

    @torch.no_grad()
    def _pack(self) -> dict:
        net, cfg = self.net, self.cfg
        # Inference (fold_weights): every re-layout and the fp64 weight folds run on the HOST, once per parameter version;
        # the device sees one upload per matrix and the multi-matrix image-pack launches — no ATen kernels in front of the
        # first compute launch.  Trainer: weights change every step, the re-layout stays on the device (captured in g_pack).
        dev = torch.device("cpu") if self.fold_weights else self.device
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)  # noqa: E731
        P: dict = {}
        tc = self.prec != lib.PREC_FP32_SIMT

        def W(bt: torch.Tensor, n: int, k: int) -> lib.Weight:
            """fp32 "in x out" matrix -> Weight (adds the tensor-core image when that engine is selected)."""
            bt = bt.to(self.device)
            return lib.Weight(bt, lib.pack_bimg(bt, n, k, transposed=True) if tc else None, n, k)

        def L(param: torch.Tensor, n: int, k: int, k_gemm: Optional[int] = None) -> lib.Weight:
            """plain nn.Linear / 1x1-conv weight [out=n, in=k].  Inference: "in x out" fp32 matrix + image.  Trainer (weights
            change every step): the image is packed straight from the parameter — no intermediate tensors, one launch —
            and the `src` field lets the data-gradient image be packed from the same storage transposed."""
            w2 = f32(param).reshape(n, k)
            if self.fold_weights or not tc:
                return W(pack_linear_t(w2, k_pad=k_gemm), n, k if k_gemm is None else k_gemm)
            kg = k if k_gemm is None else k_gemm
            assert (k + 63) // 64 == (kg + 63) // 64, "padded K must stay inside the last 64-wide k-slab"
            wobj = lib.Weight(None, lib.pack_bimg(w2, n, k, transposed=False), n, kg)
            wobj.src = w2
            return wobj

        cin_p = _ceil(cfg.in_channel, 16)
        P["cin_p"] = cin_p
        P["patch_embed"] = W(pack_conv3x3(f32(net.patch_embed.proj.weight), cin_pad=cin_p), cfg.dim, 9 * cin_p)
        P["down1_2"] = W(pack_conv3x3(f32(net.down1_2.body[0].weight)), cfg.dim // 2, 9 * cfg.dim)
        P["down2_3"] = W(pack_conv3x3(f32(net.down2_3.body[0].weight)), cfg.dim, 18 * cfg.dim)
        P["up3_2"] = W(pack_conv3x3(f32(net.up3_2.body[0].weight), shuffle=True), 8 * cfg.dim, 36 * cfg.dim)
        P["up2_1"] = W(pack_conv3x3(f32(net.up2_1.body[0].weight), shuffle=True), 4 * cfg.dim, 18 * cfg.dim)
        P["reduce_chan_level2"] = L(net.reduce_chan_level2.weight, 2 * cfg.dim, 4 * cfg.dim)
        P["output"] = W(pack_conv3x3(f32(net.output.weight)), cfg.out_channel, 18 * cfg.dim)
        if getattr(self, "_clip_dev", None) is None:  # a constant, moved once (the module keeps it on the host)
            self._clip_dev = f32(net.text_prompt.clip_prompt).contiguous()
        P["clip"] = self._clip_dev

        lane = 0
        for st in cfg.stages():
            hid = cfg.hidden(st.dim)
            hid_pad = _ceil(hid, 16)
            blocks = []
            for blk in getattr(net, st.name).blocks:
                self._pack_lane(lane)
                lane += 1
                d = {"hid_pad": hid_pad}
                d["ln1"] = (f32(blk.norm1.weight).contiguous(), f32(blk.norm1.bias).contiguous())
                d["ln2"] = (f32(blk.norm2.weight).contiguous(), f32(blk.norm2.bias).contiguous())
                a = blk.attn
                d["qkv_w"] = L(a.qkv.weight, 3 * st.dim, st.dim)
                d["qkv_b"] = f32(a.qkv.bias).contiguous()
                d["proj_w"] = L(a.proj.weight, st.dim, st.dim)
                d["proj_b"] = f32(a.proj.bias).contiguous()
                d["rpb"] = rel_pos_bias(f32(a.relative_position_bias_table), a.relative_position_index.to(dev))
                if self.fold_weights and st.dim // st.heads == 64:
                    # inference, head dim 64 (decoder_level1 / refinement: 61 % of the attention tokens): [heads, key, query]
                    # copy for the TMA-fed tcgen05 kernel (152 vs 160 us per launch at 512x512; at head dim 32 the mma.sync
                    # kernel is still ahead, 94 vs 104 us, so those stages keep it)
                    d["rpb_t"] = d["rpb"].transpose(1, 2).contiguous()
                g = blk.gobal_spectral_attn
                d["temp"] = f32(g.temperature).reshape(-1).contiguous()
                d["sqkv_w"] = L(g.qkv.weight, 3 * st.dim, st.dim)
                d["sdw"] = pack_dw(f32(g.qkv_dwconv.weight))
                d["sout_t"] = f32(g.project_out.weight).reshape(st.dim, st.dim).t().contiguous()
                l = blk.local_spectral_attn
                r = st.rank
                # fold Spatial_Attention.proj into the two matrices that consume its (window-mean) output;
                # done in fp64 so the fold itself adds no rounding beyond the final fp32 cast
                d["gate"] = {
                    "param": f32(l.prompt_param).reshape(PROMPT_LEN, r).contiguous(),
                    "qT": f32(l.q.weight).t().contiguous(),
                    "kvT": f32(l.kv.weight).t().contiguous(),
                    "p2T": f32(l.proj.weight).t().contiguous(),
                    "p2b": f32(l.proj.bias).contiguous(),
                    "upT": f32(l.linear_up.weight).t().contiguous(),
                }
                if not self.fold_weights:
                    # training: weights change every step, so the fp64 folds below are not worth re-doing; the local gate is
                    # computed from the window mean of the *projected* attention output with the plain matrices instead
                    nlp = _ceil(PROMPT_LEN + r, 8)
                    cat = f32(l.linear_prompt.weight).new_zeros(nlp, st.dim)
                    cat[:PROMPT_LEN] = f32(l.linear_prompt.weight)
                    cat[PROMPT_LEN:PROMPT_LEN + r] = f32(l.linear_down.weight)
                    d["ll_w"] = W(pack_linear_t(cat), nlp, st.dim)
                    fc1_w, d["fc1_b"] = pack_glu_fc1(f32(blk.mlp.fc1.weight), f32(blk.mlp.fc1.bias), hid, hid_pad)
                    d["fc1_w"] = W(fc1_w, 2 * hid_pad, st.dim)
                    d["fc2_w"] = L(blk.mlp.fc2.weight, st.dim, hid, k_gemm=hid_pad)
                    d["fc2_b"] = f32(blk.mlp.fc2.bias).contiguous()
                    blocks.append(d)
                    continue
                pw64, pb64 = a.proj.weight.detach().to(dev).double(), a.proj.bias.detach().to(dev).double()
                lp64, ld64 = l.linear_prompt.weight.detach().to(dev).double(), l.linear_down.weight.detach().to(dev).double()
                if tc and st.dim % 32 == 0:
                    # fold proj in front of the global-spectral qkv 1x1 as well (:216 then :101): one GEMM over the
                    # window-attention output with rows [W_p ; W_sqkv W_p], bias [b_p ; W_sqkv b_p]  (MPHSIR_EPI_PROJ)
                    sq64 = g.qkv.weight.detach().to(dev).double().reshape(3 * st.dim, st.dim)
                    d["projf_w"] = W(pack_linear_t(torch.cat([pw64, sq64 @ pw64], 0).float()), 4 * st.dim, st.dim)
                    d["projf_b"] = torch.cat([pb64, sq64 @ pb64]).float().contiguous()
                # prompt logits and low-rank projection of the window mean as one GEMM: rows [W_prompt W_p ; W_down W_p]
                d["gate_cat_w"] = W(pack_linear_t(torch.cat([lp64 @ pw64, ld64 @ pw64], 0).float()), PROMPT_LEN + r, st.dim)
                d["gate_cat_b"] = torch.cat([lp64 @ pb64, ld64 @ pb64]).float().contiguous()
                d["gate"].update({
                    "promptT": (lp64 @ pw64).t().float().contiguous(),
                    "promptb": (lp64 @ pb64).float().contiguous(),
                    "downT": (ld64 @ pw64).t().float().contiguous(),
                    "downb": (ld64 @ pb64).float().contiguous(),
                })
                fc1_w, d["fc1_b"] = pack_glu_fc1(f32(blk.mlp.fc1.weight), f32(blk.mlp.fc1.bias), hid, hid_pad)
                d["fc1_w"] = W(fc1_w, 2 * hid_pad, st.dim)
                d["fc2_w"] = W(pack_linear_t(f32(blk.mlp.fc2.weight), k_pad=hid_pad), st.dim, hid_pad)
                d["fc2_b"] = f32(blk.mlp.fc2.bias).contiguous()
                blocks.append(d)
            P[st.name] = blocks

        def ln(m):
            return (f32(m.body.weight).contiguous(), f32(m.body.bias).contiguous())

        for name in ("prompt1", "prompt2"):
            self._pack_lane(lane)
            lane += 1
            m = getattr(net, name)
            D = m.visual_prompt.shape[1]
            ps = m.prompt_size
            hid = cfg.hidden(D)
            hid_pad = _ceil(hid, 16)
            ct = m.cross_transformer
            d = {"D": D, "ps": ps, "hid_pad": hid_pad}
            d["learnable"] = f32(m.text_prompt_learnable).reshape(cfg.task_classes, D).contiguous()
            d["visual"] = f32(m.visual_prompt)[0].permute(1, 2, 0).reshape(ps * ps, D).contiguous()
            d["ln11"], d["ln12"], d["ln2"] = ln(ct.norm11), ln(ct.norm12), ln(ct.norm2)
            d["q_w"] = L(ct.attn.q.weight, D, D)
            d["q_dw"] = pack_dw(f32(ct.attn.q_dwconv.weight))
            d["kv_w"] = L(ct.attn.kv.weight, 2 * D, D)
            d["kv_dw"] = pack_dw(f32(ct.attn.kv_dwconv.weight))
            d["temp"] = f32(ct.attn.temperature).reshape(-1).contiguous()
            d["out_t"] = f32(ct.attn.project_out.weight).reshape(D, D).t().contiguous()
            pin_w, d["ffn_dw"], pout_w = pack_gdfn(
                f32(ct.ffn.project_in.weight), f32(ct.ffn.dwconv.weight), f32(ct.ffn.project_out.weight), hid, hid_pad)
            d["pin_w"], d["pout_w"] = W(pin_w, 2 * hid_pad, D), W(pout_w, D, hid_pad)
            d["conv_last"] = W(pack_conv3x3(f32(m.conv_last.weight)), D, 9 * D)
            P[name] = d

        for name in ("fusion1", "fusion2"):
            self._pack_lane(lane)
            lane += 1
            m = getattr(net, name)
            tb = m.transformer
            C2 = tb.norm1.body.weight.shape[0]
            hid = cfg.hidden(C2)
            hid_pad = _ceil(hid, 16)
            d = {"C": C2, "heads": m.heads, "hid_pad": hid_pad}
            d["ln1"], d["ln2"] = ln(tb.norm1), ln(tb.norm2)
            d["qkv_w"] = L(tb.attn.qkv.weight, 3 * C2, C2)
            d["dw"] = pack_dw(f32(tb.attn.qkv_dwconv.weight))
            d["temp"] = f32(tb.attn.temperature).reshape(-1).contiguous()
            d["out_t"] = f32(tb.attn.project_out.weight).reshape(C2, C2).t().contiguous()
            pin_w, d["ffn_dw"], pout_w = pack_gdfn(
                f32(tb.ffn.project_in.weight), f32(tb.ffn.dwconv.weight), f32(tb.ffn.project_out.weight), hid, hid_pad)
            d["pin_w"], d["pout_w"] = W(pin_w, 2 * hid_pad, C2), W(pout_w, C2, hid_pad)
            d["conv_w"] = L(m.conv.weight, C2 // 2, C2)
            P[name] = d
        self._pack_lane(-1)
        return self._to_device(P) if dev.type == "cpu" else P

    def _to_device(self, obj):
        """move the host-packed leaves (LayerNorm vectors, biases, tap tables, ...) to the device"""
        if isinstance(obj, torch.Tensor):
            return obj.to(self.device)
        if isinstance(obj, dict):
            return {k: self._to_device(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to_device(v) for v in obj)
        return obj

    # -- building blocks -------------------------------------------------------------------------
    def _gemm(self, *a, **k):
        lib.gemm(*a, precision=self.prec, **k)

    def _conv(self, *a, **k):
        lib.conv3x3(*a, precision=self.prec, **k)

    def _spectral_attention(self, tag: str, q: View, q_shared: bool, k: View, k_shared: bool, temp, out_t,
                            B: int, HW: int, heads: int, c: int) -> lib.Weight:
        """Gram statistics -> softmax -> folded per-sample matrix Mt [B, C, ldm] ("in x out")."""
        ws = self.ws
        nfl, nch = lib.gram_partial_floats(B, heads, c, HW)
        partial = ws.flat(tag + ".partial", nfl)
        lib.gram_partial(q, q_shared, k, k_shared, partial, B, HW, heads, c)
        return self._spectral_finish(tag, partial, nch, temp, out_t, B, heads, c)

    def _spectral_finish(self, tag: str, partial, nch: int, temp, out_t, B: int, heads: int, c: int) -> lib.Weight:
        ws = self.ws
        C = heads * c
        scratch = ws.flat(tag + ".gsum", B * heads * (c * c + 2 * c))
        Mt = img = None
        if self.prec == lib.PREC_FP32_SIMT:
            ldm = _ldb(C)
            Mt = ws.flat(tag + ".Mt", B * _ceil(C, 16) * ldm).view(-1)[: B * _ceil(C, 16) * ldm].view(B, _ceil(C, 16), ldm)
        else:
            nb = lib.bimg_bytes(C, C)
            img = ws.raw(tag + ".img", B * nb)[: B * nb].view(B, nb)
        lib.spectral_finish(partial, nch, scratch, temp, out_t, Mt, img, B, heads, c)
        return lib.Weight(Mt, img, C, C)

    def _global_spectral(self, tag: str, t3: View, w_dw, temp, out_t, B: int, H: int, W: int, C: int, heads: int):
        """qkv_dwconv + Gram + softmax + fold (net/MP_HSIR.py:98-113): returns (v operand view, folded Weight).
        Tensor-core precisions use the fused dwconv+Gram kernel (q, k never reach HBM)."""
        ws = self.ws
        N = B * H * W
        c = C // heads
        if self.band is not None:
            return self._global_spectral_band(tag, t3, w_dw, temp, out_t, H, W, C, heads)
        if self.prec != lib.PREC_FP32_SIMT and lib.dwgram_supported(C, c):
            v = ws.mat(tag + ".v", N, C)
            nfl, nch = lib.dwgram_partial_floats(B, heads, c, H, W)
            partial = ws.flat(tag + ".partial", nfl)
            lib.dwgram(t3, w_dw, v, partial, B, H, W, C, heads, self.prec)
            return v, self._spectral_finish(tag, partial, nch, temp, out_t, B, heads, c)
        dw3 = ws.mat(tag + ".dw3", N, 3 * C)
        lib.dwconv3x3(t3, w_dw, dw3, B, H, W, 3 * C)
        Mt = self._spectral_attention(tag, dw3.cols_slice(0, C), False, dw3.cols_slice(C, 2 * C), False, temp, out_t,
                                      B, H * W, heads, c)
        return dw3.cols_slice(2 * C, 3 * C), Mt

    def _global_spectral_band(self, tag: str, t3: View, w_dw, temp, out_t, H: int, W: int, C: int, heads: int):
        """`_global_spectral` on a row band: depthwise conv + V on the rows [a, b) this rank may touch (own rows + halo; at
        the scene's top / bottom the view ends at the scene edge, so the kernels' zero padding IS the conv padding), Gram
        statistics on the rank's OWN rows only, summed over the ranks (ONE small all-reduce per block: heads*(c*c+2c)
        floats), then the softmax / fold on every rank (net/MP_HSIR.py:104-113 reduce over the whole scene)."""
        ws, bd = self.ws, self.band
        c = C // heads
        a, b = bd.a, bd.b
        v = ws.mat(tag + ".v", H * W, C)
        t3v, vv = t3.rows_slice(a * W, b * W), v.rows_slice(a * W, b * W)
        own0 = bd.halo - a
        if lib.dwgram_supported(C, c):
            nfl, nch = lib.dwgram_partial_floats(1, heads, c, b - a, W)
            partial = ws.flat(tag + ".partial", nfl)
            lib.dwgram(t3v, w_dw, vv, partial, 1, b - a, W, C, heads, self.prec, gram_rows=(own0, own0 + bd.Hb))
            vout = v
        else:
            dw3 = ws.mat(tag + ".dw3", H * W, 3 * C)
            lib.dwconv3x3(t3v, w_dw, dw3.rows_slice(a * W, b * W), 1, b - a, W, 3 * C)
            own = dw3.rows_slice(bd.halo * W, (bd.halo + bd.Hb) * W)
            nfl, nch = lib.gram_partial_floats(1, heads, c, bd.Hb * W)
            partial = ws.flat(tag + ".partial", nfl)
            lib.gram_partial(own.cols_slice(0, C), False, own.cols_slice(C, 2 * C), False, partial, 1, bd.Hb * W, heads, c)
            vout = dw3.cols_slice(2 * C, 3 * C)
        per = heads * (c * c + 2 * c)
        gsum = ws.flat(tag + ".gsum_rank", per)
        lib.gram_reduce(partial, nch, gsum, 1, heads, c)
        bd.comm.all_reduce(gsum[:per])
        return vout, self._spectral_finish(tag, gsum, 1, temp, out_t, 1, heads, c)

    def _pgsstb(self, w: dict, st: Stage, shift: int, x: View, out: View, res2: Optional[View], B: int, H: int,
                W: int, row_scales=None, taps: Optional[dict] = None):
        ws = self.ws
        C, heads = st.dim, st.heads
        N = B * H * W
        B_ = N // 64
        qkv = ws.mat("qkv", N, 3 * C)
        core = ws.mat("core", N, C)
        sa = ws.mat("sa", N, C)
        mid = ws.mat("mid", N, C)
        wmean = ws.flat("wmean", B_ * C)
        gate = ws.flat("gate", B_ * C)
        s1 = None if row_scales is None else row_scales[0]
        s2 = None if row_scales is None else row_scales[1]
        mask_H, mask_y0 = None, 0
        if self.band is not None:
            # row band: refresh the 8-row halos of the block input from the neighbour ranks (cyclic: the roll of :672 wraps
            # between the last and the first band); everything below is then local except the Gram all-reduce
            self.band.exchange(x)
            mask_H, mask_y0 = self.band.Hg, self.band.y0

        # LN1 + qkv projection (net/MP_HSIR.py:667, :195)
        self._gemm(x, w["qkv_w"], qkv, 3 * C, ln=w["ln1"], bias=w["qkv_b"])
        # shifted-window attention core + per-window mean (:671-683, :198-215)
        lib.window_attn(qkv, w["rpb"], core, wmean, B, H, W, C, heads, shift, precision=self.prec, mask_H=mask_H, mask_y0=mask_y0,
                        bias_t=w.get("rpb_t"))
        # local spectral gate (:132-152)
        if "gate_cat_w" not in w:
            pass  # un-folded weights (trainer): the gate is computed below from the window mean of sa
        elif self.fused_gate and st.rank % 4 == 0:
            lib.local_gate2(wmean, w["gate"], gate, B_, C, st.rank)
        elif self.split_gate and st.rank % 4 == 0:
            logits = ws.mat("gate_logits", B_, _ceil(PROMPT_LEN + st.rank, 16))
            self._gemm(View(wmean.data_ptr(), C, B_, C, wmean), w["gate_cat_w"], logits, PROMPT_LEN + st.rank, bias=w["gate_cat_b"])
            lib.local_gate_tail(logits, w["gate"], gate, B_, C, st.rank)
        else:
            lib.local_gate(wmean, w["gate"], gate, B_, C, st.rank)
        t3 = ws.mat("qkv", N, 3 * C)  # qkv is dead: reuse
        if "projf_w" in w and self.fold_proj and taps is None:
            # proj (:216) and the global-spectral qkv 1x1 (:101) as ONE GEMM over core; its left C columns leave as
            # u = shortcut + DropPath(sa*gate) (:715-718, :153), so sa itself is never materialised
            self._gemm(core, w["projf_w"], sa, 4 * C, bias=w["projf_b"], epi=lib.EPI_PROJ, res1=x, gate=gate, Y2=t3,
                       n_split=C, H=H, W=W, shift=shift, rows_per_batch=H * W, row_scale=s1)
            v, Mt = self._global_spectral("spec", t3, w["sdw"], w["temp"], w["sout_t"], B, H, W, C, heads)
            # mid = u + DropPath(project_out(attn v))
            self._gemm(v, Mt, mid, C, epi=lib.EPI_RESIDUAL, res1=sa, rows_per_batch=H * W, row_scale=s1)
        else:
            # attention output projection (:216) in image order
            self._gemm(core, w["proj_w"], sa, C, bias=w["proj_b"])
            if "gate_cat_w" not in w:
                self._gate_from_sa(w, st, sa, gate, B, H, W, shift)
            # global spectral attention: 1x1 -> dw3x3 -> Gram/softmax/fold -> apply (:98-113)
            self._gemm(sa, w["sqkv_w"], t3, 3 * C)
            v, Mt = self._global_spectral("spec", t3, w["sdw"], w["temp"], w["sout_t"], B, H, W, C, heads)
            # x = shortcut + DropPath(sa*gate + project_out(attn v))   (:715-718)
            self._gemm(v, Mt, mid, C, epi=lib.EPI_SPECTRAL, res1=x, gsrc=sa, gate=gate,
                       H=H, W=W, shift=shift, rows_per_batch=H * W, row_scale=s1)
        # x = x + DropPath(fc2(value * gelu(gate)))  with LN2 fused in front (:719, :76-82)
        if self.prec != lib.PREC_FP32_SIMT and self.fuse_mlp and lib.mlp_supported(C, w["hid_pad"]):
            lib.mlp(mid, w["ln2"], w["fc1_w"], w["fc1_b"], w["fc2_w"], w["fc2_b"], out, w["hid_pad"], self.prec,
                    res2=res2, row_scale=s2, rows_per_batch=H * W)
        else:
            hidden = ws.mat("hidden", N, w["hid_pad"])
            self._gemm(mid, w["fc1_w"], hidden, 2 * w["hid_pad"], ln=w["ln2"], bias=w["fc1_b"], epi=lib.EPI_GLU)
            self._gemm(hidden, w["fc2_w"], out, C, bias=w["fc2_b"], epi=lib.EPI_RESIDUAL, res1=mid, res2=res2,
                       rows_per_batch=H * W, row_scale=s2)
        if taps is not None:
            taps.update(core=core.torch().clone(), wmean=wmean[: B_ * C].view(B_, C).clone(),
                        gate=gate[: B_ * C].view(B_, C).clone(), sa=sa.torch().clone(), mid=mid.torch().clone(),
                        out=out.torch().clone())

    def _gate_from_sa(self, w: dict, st: Stage, sa: View, gate, B: int, H: int, W: int, shift: int, msa=None, LL=None):
        """local spectral gate from the window mean of the projected attention output (:135-152), plain weights"""
        C, B_ = st.dim, B * H * W // 64
        msa = self.ws.flat("gate_msa", B_ * C) if msa is None else msa
        lib.window_reduce(sa, None, msa, B, H, W, C, shift, 1.0 / 64.0)
        nl = w["ll_w"].n
        LL = self.ws.mat("gate_LL", B_, _ceil(nl, 16)) if LL is None else LL
        self._gemm(View(msa.data_ptr(), C, B_, C, msa), w["ll_w"], LL, nl)
        lib.local_gate_tail(LL, w["gate"], gate, B_, C, st.rank)

    def _stage(self, name: str, x_in: View, out: View, B: int, H: int, W: int):
        st = {s.name: s for s in self.cfg.stages()}[name]
        blocks = self.packed[name]
        N = B * H * W
        ping = [self.ws.mat("xping", N, st.dim), self.ws.mat("xpong", N, st.dim)]
        x = x_in
        for i, w in enumerate(blocks):
            last = i == len(blocks) - 1
            dst = out if last else ping[i % 2]
            # BaseBlock residual `x + shortcut` (net/MP_HSIR.py:760) rides on the last fc2 epilogue
            self._pgsstb(w, st, SHIFT if i % 2 else 0, x, dst, x_in if last else None, B, H, W)
            x = dst

    def _gdfn(self, tag: str, w: dict, x: View, out: View, ln, B: int, H: int, W: int, D: int):
        """x + GDFN(LN(x)) (net/MP_HSIR.py:477, :286, :386-391)."""
        N = B * H * W
        hp = w["hid_pad"]
        hin = self.ws.mat(tag + ".hin", N, 2 * hp)
        hg = self.ws.mat(tag + ".hg", N, hp)
        self._gemm(x, w["pin_w"], hin, 2 * hp, ln=ln)
        if self.band is not None:   # the rows this rank may touch; the view ends at the scene edge on the outer ranks
            a, b = self.band.a, self.band.b
            lib.dwconv3x3(hin.rows_slice(a * W, b * W), w["ffn_dw"], hg.rows_slice(a * W, b * W), 1, b - a, W, 2 * hp, gate_half=hp)
        else:
            lib.dwconv3x3(hin, w["ffn_dw"], hg, B, H, W, 2 * hp, gate_half=hp)
        self._gemm(hg, w["pout_w"], out, D, epi=lib.EPI_RESIDUAL, res1=x)

    def _tvsp_key(self, task_key, B: int, Hs: int, Ws: int, out: View):
        return None if task_key is None else (task_key, B, Hs, Ws, out.ptr, out.ld, self._pack_serial, self.ws.generation)

    def _tvsp_cached(self, name: str, clip_b, weights, B: int, Hs: int, Ws: int, out: View, task_key) -> None:
        """`_tvsp` unless `out` already holds the prompt of these task ids at this shape (inference-time cache)."""
        key = self._tvsp_key(task_key, B, Hs, Ws, out)
        if key is not None and self._tvsp_valid.get(name) == key:
            return
        self._tvsp(name, clip_b, weights, B, Hs, Ws, out)
        self._tvsp_valid[name] = key

    def _tvsp(self, name: str, clip_b: torch.Tensor, weights: torch.Tensor, B: int, Hs: int, Ws: int, out: View):
        """TVSP.forward (net/MP_HSIR.py:572-583) -> writes [B*Hs*Ws, D] into `out` (a column slice)."""
        w = self.packed[name]
        ws = self.ws
        D, ps = w["D"], w["ps"]
        T = self.cfg.task_classes
        n = ps * ps
        Q = ws.mat(name + ".Q", B * n, D)
        lib.tvsp_query(clip_b, weights, w["learnable"], Q, B, T, D, ps)
        q1 = ws.mat(name + ".q1", B * n, D)
        qd = ws.mat(name + ".qd", B * n, D)
        self._gemm(Q, w["q_w"], q1, D, ln=w["ln11"])
        lib.dwconv3x3(q1, w["q_dw"], qd, B, ps, ps, D)
        vis = View.of(w["visual"])
        kv1 = ws.mat(name + ".kv1", n, 2 * D)
        kvd = ws.mat(name + ".kvd", n, 2 * D)
        self._gemm(vis, w["kv_w"], kv1, 2 * D, ln=w["ln12"])
        lib.dwconv3x3(kv1, w["kv_dw"], kvd, 1, ps, ps, 2 * D)
        Mt = self._spectral_attention(name + ".spec", qd, False, kvd.cols_slice(0, D), True, w["temp"], w["out_t"],
                                      B, n, 2, D // 2)
        xa = ws.mat(name + ".xa", B * n, D)
        self._gemm(kvd.cols_slice(D, 2 * D), Mt, xa, D, epi=lib.EPI_RESIDUAL, res1=Q, rows_per_batch=n, a_row_mod=n, M=B * n)
        pr = ws.mat(name + ".pr", B * n, D)
        self._gdfn(name + ".ffn", w, xa, pr, w["ln2"], B, ps, ps, D)
        if (Hs, Ws) != (ps, ps):
            prr = ws.mat(name + ".prr", B * Hs * Ws, D)
            lib.bilinear(pr, prr, B, ps, ps, Hs, Ws, D)
            pr = prr
        self._conv(pr, w["conv_last"], out.ptr, out.ld, B, Hs, Ws, D, D, lib.CONV_TOKENS)

    def _fusion(self, name: str, xcat: View, out: View, B: int, H: int, W: int):
        """PromptFusion.forward (net/MP_HSIR.py:594-599) on the already-concatenated buffer."""
        w = self.packed[name]
        ws = self.ws
        C2, heads = w["C"], w["heads"]
        N = B * H * W
        t3 = ws.mat(name + ".t3", N, 3 * C2)
        self._gemm(xcat, w["qkv_w"], t3, 3 * C2, ln=w["ln1"])
        v, Mt = self._global_spectral(name + ".spec", t3, w["dw"], w["temp"], w["out_t"], B, H, W, C2, heads)
        y1 = ws.mat(name + ".y1", N, C2)
        self._gemm(v, Mt, y1, C2, epi=lib.EPI_RESIDUAL, res1=xcat, rows_per_batch=H * W)
        y2 = ws.mat(name + ".y2", N, C2)
        self._gdfn(name + ".ffn", w, y1, y2, w["ln2"], B, H, W, C2)
        self._gemm(y2, w["conv_w"], out, out.cols)

    # -- whole network ----------------------------------------------------------------------------
    def task_weights(self, task_id: torch.Tensor) -> torch.Tensor:
        """prompt_weights of Text_Prompt.forward (net/MP_HSIR.py:519-525) as fp32 [B,T]."""
        T = self.cfg.task_classes
        tid = task_id.to(self.device).long()
        w = F.one_hot(tid, T).to(torch.float32)
        if tid.dim() > 1:
            w = w.mean(dim=1)
        return w.contiguous()

    @torch.no_grad()
    def forward(self, inp: torch.Tensor, task_id: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        if inp.dim() != 4 or inp.shape[1] != cfg.in_channel:
            raise ValueError(f"expected input [B,{cfg.in_channel},H,W], got {tuple(inp.shape)}")
        B, _, H, W = inp.shape
        if H % 32 or W % 32:
            raise ValueError(f"H and W must be multiples of 32 (two 2x down-samplings x window 8), got {H}x{W}")
        if task_id.shape[0] != B:
            raise ValueError("task_id batch dimension does not match the input")
        if inp.device != self.device:
            raise RuntimeError(f"input on {inp.device} but parameters on {self.device}")
        if cfg.in_channel != cfg.out_channel:
            raise ValueError("global residual requires in_channel == out_channel (net/MP_HSIR.py:841)")
        self._ensure_packed()
        x = inp.detach().to(torch.float32).contiguous()
        # host copy of the task ids = the key of the prompt cache (a device tensor costs one tiny synchronising read, as
        # test.py's own loop does for its metrics; pass a host tensor to avoid it)
        task_key = None
        if self.cache_prompts:
            task_key = (tuple(task_id.shape), tuple(task_id.reshape(-1).tolist()))
        with torch.cuda.device(self.device):
            weights = self.task_weights(task_id)
            if self.net.use_cuda_graph and lib.PROFILER is None:
                out = self._run_graphed(x, weights, task_key)
            else:
                out = torch.empty_like(x)
                self._run(x, weights, out, task_key=task_key)
        return out.to(inp.dtype)

    def _run_graphed(self, x: torch.Tensor, weights: torch.Tensor, task_key=None) -> torch.Tensor:
        """Replay the whole forward as one CUDA graph per (input shape, task ids) (every launch goes through the
        C ABI on torch's capture stream; all buffers are workspace-owned, so pointers are stable).
        A workspace re-allocation (a larger shape came along) drops every captured graph.  With the prompt cache the
        captured graph holds no TVSP launches: it reads the prompt columns, which are refreshed eagerly before a
        replay whenever another call overwrote them."""
        key = (tuple(x.shape), task_key)
        ent = self._graphs.get(key)
        if ent is not None and ent[0] == self.ws.generation and task_key is not None and ent[6] != self._tvsp_valid:
            self._refresh_prompts(x.shape, weights, task_key)
        if ent is None or ent[0] != self.ws.generation:
            sx, sw, so = torch.empty_like(x), torch.empty_like(weights), torch.empty_like(x)
            sx.copy_(x)
            sw.copy_(weights)
            self._run(sx, sw, so, task_key=task_key)  # eager: sizes the workspace, sets kernel attributes, fills the prompt cache
            torch.cuda.current_stream().synchronize()
            gen = self.ws.generation
            for k in [k for k, e in self._graphs.items() if e[0] != gen]:
                del self._graphs[k]
            g = torch.cuda.CUDAGraph()
            n0 = lib.LAUNCHES
            with torch.cuda.graph(g):
                self._run(sx, sw, so, task_key=task_key)
            if self.ws.generation != gen:
                raise RuntimeError("workspace grew during graph capture")
            ent = (gen, g, sx, sw, so, lib.LAUNCHES - n0, dict(self._tvsp_valid))
            lib.LAUNCHES = n0  # capture records, it does not launch
            self._graphs[key] = ent
        _, g, sx, sw, so, n_kernels, _ = ent
        sx.copy_(x)
        sw.copy_(weights)
        g.replay()
        lib.LAUNCHES += n_kernels
        return so.clone()

    def _refresh_prompts(self, shape, weights: torch.Tensor, task_key) -> None:
        """re-run text prompt + both TVSPs into the prompt columns (a call with other task ids / another shape overwrote them)"""
        cfg, P, ws = self.cfg, self.packed, self.ws
        B, _, H, W = shape
        d = cfg.dim
        clip_b = ws.flat("clip_b", B * 512)
        lib.text_prompt(weights, P["clip"], clip_b, B, cfg.task_classes)
        fcat1 = ws.mat("fcat1", B * H * W, 2 * d)
        fcat2 = ws.mat("fcat2", B * H * W // 4, 4 * d)
        self._tvsp_cached("prompt2", clip_b, weights, B, H // 2, W // 2, fcat2.cols_slice(2 * d, 4 * d), task_key)
        self._tvsp_cached("prompt1", clip_b, weights, B, H, W, fcat1.cols_slice(d, 2 * d), task_key)

    def _run(self, x: torch.Tensor, weights: torch.Tensor, out: torch.Tensor, taps: Optional[dict] = None, task_key=None):
        cfg, P, ws = self.cfg, self.packed, self.ws
        B, _, H, W = x.shape
        d = cfg.dim
        N1, N2, N3 = B * H * W, B * H * W // 4, B * H * W // 16
        H2, W2, H3, W3 = H // 2, W // 2, H // 4, W // 4
        T = cfg.task_classes

        clip_b = ws.flat("clip_b", B * 512)
        lib.text_prompt(weights, P["clip"], clip_b, B, T)

        tok = ws.mat("tok_in", N1, P["cin_p"])
        lib.nchw_to_tokens(x, tok)
        x1 = ws.mat("x1", N1, d)
        self._conv(tok, P["patch_embed"], x1.ptr, x1.ld, B, H, W, P["cin_p"], d)

        fcat1 = ws.mat("fcat1", N1, 2 * d)          # [e1 | prompt1]
        e1 = fcat1.cols_slice(0, d)
        self._stage("encoder_level1", x1, e1, B, H, W)

        x2 = ws.mat("x2", N2, 2 * d)
        self._conv(e1, P["down1_2"], x2.ptr, x2.ld, B, H, W, d, d // 2, lib.CONV_UNSHUFFLE)
        fcat2 = ws.mat("fcat2", N2, 4 * d)          # [e2 | prompt2]
        e2 = fcat2.cols_slice(0, 2 * d)
        self._stage("encoder_level2", x2, e2, B, H2, W2)

        x3 = ws.mat("x3", N3, 4 * d)
        self._conv(e2, P["down2_3"], x3.ptr, x3.ld, B, H2, W2, 2 * d, d, lib.CONV_UNSHUFFLE)
        lat = ws.mat("lat", N3, 4 * d)
        self._stage("latent", x3, lat, B, H3, W3)

        cat2 = ws.mat("cat2", N2, 4 * d)            # [up3_2(latent) | fusion2]
        self._conv(lat, P["up3_2"], cat2.ptr, cat2.ld, B, H3, W3, 4 * d, 8 * d, lib.CONV_SHUFFLE)
        self._tvsp_cached("prompt2", clip_b, weights, B, H2, W2, fcat2.cols_slice(2 * d, 4 * d), task_key)
        self._fusion("fusion2", fcat2, cat2.cols_slice(2 * d, 4 * d), B, H2, W2)
        d2in = ws.mat("d2in", N2, 2 * d)
        self._gemm(cat2, P["reduce_chan_level2"], d2in, 2 * d)
        d2 = ws.mat("d2", N2, 2 * d)
        self._stage("decoder_level2", d2in, d2, B, H2, W2)

        cat1 = ws.mat("cat1", N1, 2 * d)            # [up2_1(d2) | fusion1]
        self._conv(d2, P["up2_1"], cat1.ptr, cat1.ld, B, H2, W2, 2 * d, 4 * d, lib.CONV_SHUFFLE)
        self._tvsp_cached("prompt1", clip_b, weights, B, H, W, fcat1.cols_slice(d, 2 * d), task_key)
        self._fusion("fusion1", fcat1, cat1.cols_slice(d, 2 * d), B, H, W)
        dd1 = ws.mat("dd1", N1, 2 * d)
        self._stage("decoder_level1", cat1, dd1, B, H, W)
        ref = ws.mat("ref", N1, 2 * d)
        self._stage("refinement", dd1, ref, B, H, W)

        self._conv(ref, P["output"], out.data_ptr(), 0, B, H, W, 2 * d, cfg.out_channel, lib.CONV_NCHW_RES, R=x)
        if taps is not None:
            taps.update(x1=x1.torch().clone(), e1=e1.torch().clone(), e2=e2.torch().clone(), lat=lat.torch().clone(),
                        p1=fcat1.cols_slice(d, 2 * d).torch().clone(), p2=fcat2.cols_slice(2 * d, 4 * d).torch().clone(),
                        f1=cat1.cols_slice(d, 2 * d).torch().clone(), f2=cat2.cols_slice(2 * d, 4 * d).torch().clone(),
                        d1=ref.torch().clone())
