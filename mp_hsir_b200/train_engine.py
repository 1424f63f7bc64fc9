"""Training executor: forward-with-saved-activations, hand-written backward and AdamW for MP_HSIR_Net.

Replaces what autograd + torch.optim.AdamW + DDP do for ``train.py:50-69,118`` of the reference: one
``train_step(degraded, clean, task_id)`` = forward, ``L1(clamp(out,0,1), clean)`` (train.py:58-61), the
backward pass of every op of ``net/MP_HSIR.py`` and one AdamW update.  Host logic only: every arithmetic
step is a libmphsir.so launch (``lib.py``); torch allocates memory, zeroes buffers, re-lays-out *weights*
and (multi-GPU) all-reduces the flat gradient buffer over NCCL.

Layout choices
  * parameters live in ONE flat fp32 buffer (each ``nn.Parameter`` is re-pointed to a 16-byte aligned view
    of it), gradients in a parallel flat buffer (``param.grad`` are views): AdamW is one launch, the DDP
    all-reduce is one collective, ``zero_grad`` one memset.
  * the 8 parameters the reference never uses (``prompt{1,2}.{text,clip}_linear``, net/MP_HSIR.py:552,555;
    ``grad is None`` there, so torch's AdamW skips them) sit outside the flat range and stay untouched.
  * activations needed by the backward are kept per block in named workspace buffers (180 GB HBM: a batch
    of 32 64x64 patches keeps ~16 GB).  Two things are recomputed instead of stored because their forward
    kernels are fused: the MLP's hidden tile (LN2 + fc1 GEMM again) and the depthwise-conv outputs q, k of the
    global spectral attention (the forward runs the fused dwconv+Gram kernel that writes only v).
  * data gradients of linear / conv layers run on the forward's tcgen05 GEMM engine with transposed
    (flipped) weight images; weight gradients (and the bias gradients riding on them) on ``mphsir_wgrad``:
    a tcgen05 kernel for convs / wide layers, an mma.sync kernel for narrow ones, one multi-problem launch
    for the r-sized local-gate matrices.
  * after every optimizer step the weight images are re-packed from the flat parameter buffer: plain linear
    layers straight from the parameter storage, all images through a handful of multi-matrix launches.
"""
from __future__ import annotations

import os

from typing import Dict, Optional

import torch

from . import lib
from .config import PROMPT_LEN, SHIFT, Stage
from .engine import Engine, _ceil, _ldb, pack_conv3x3
from .lib import MAP_HALVES, MAP_INTERLEAVE, View

DEAD_PARAMS = ("text_linear", "clip_linear")  # defined, never used by TVSP.forward (net/MP_HSIR.py:572-583)


def _flip_dw(w9: torch.Tensor) -> torch.Tensor:
    """[9, C] tap-major depthwise weights -> the data-gradient kernel (taps reversed)."""
    return w9.flip(0).contiguous()


class TrainEngine(Engine):
    def __init__(self, net, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(net)
        if self.prec == lib.PREC_FP32_SIMT:
            raise ValueError("training runs on the tensor-core engine: precision must be 'fp32' (bf16x3) or 'bf16'")
        self.fold_weights = False
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self._flatten_parameters()
        self.loss_buf = torch.zeros(1, device=self.device, dtype=torch.float32)
        self._train_packed_version = None
        # streams the captured weight re-pack forks into (0 = one chain); env switch for A/B runs
        self.pack_lanes = int(os.environ.get("MPHSIR_PACK_LANES", "13"))
        rates = [1.0 - r for st in self.cfg.stages() for r in st.dpr if r > 0.0]   # DropPath keep probabilities, block order
        self._dp_keep_prob = torch.tensor(rates, device=self.device, dtype=torch.float32).view(-1, 1, 1) if rates else None
        self.time_graphs = False   # bench: CUDA events around the three graphs of the captured step -> self.graph_ms
        self.graph_ms = None

    # -- flat parameter / gradient storage -----------------------------------------------------------
    def _flatten_parameters(self):
        named = [(n, p) for n, p in self.net.named_parameters() if not any(d in n for d in DEAD_PARAMS)]
        off, offsets = 0, {}
        for n, p in named:
            offsets[n] = off
            off += _ceil(p.numel(), 4)
        self.flat_p = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.flat_g = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.flat_m = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.flat_v = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.g: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for n, p in named:
                o, k = offsets[n], p.numel()
                self.flat_p[o:o + k].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = self.flat_p[o:o + k].view(p.shape)
                self.g[n] = self.flat_g[o:o + k]
                p.grad = self.flat_g[o:o + k].view(p.shape)
        self.param_offsets = offsets
        self.invalidate()

    def zero_grad(self):
        self.flat_g.zero_()

    # -- weights ---------------------------------------------------------------------------------------
    def _W(self, bt: torch.Tensor, n: int, k: int) -> lib.Weight:
        return lib.Weight(bt, lib.pack_bimg(bt, n, k, transposed=True), n, k)

    def _dgrad_of(self, w: lib.Weight) -> lib.Weight:
        """data-gradient weight of a forward GEMM weight: dX[M,K_f] = dY[M,N_f] @ W  (packed column order kept)."""
        n_f, k_f = w.n, w.k
        if w.src is not None:
            # packed straight from the parameter [n_f, k_valid]: the data-gradient image is the same storage read transposed
            # (logical [N = k_valid, K = n_f]); the GEMM's N = k_f may be the 16-padded k_valid, which the image rows cover
            kv = w.src.shape[1]
            assert _ceil(kv, 16) == _ceil(k_f, 16)
            return lib.Weight(None, lib.pack_bimg(w.src, kv, n_f, transposed=True), k_f, n_f)
        out = w.bt.new_zeros(_ceil(n_f, 16), _ldb(k_f))
        out[:n_f, :k_f] = w.bt[:k_f, :n_f].t()
        return self._W(out.contiguous(), k_f, n_f)

    def _conv_dgrad(self, weight: torch.Tensor, cout_pad: Optional[int] = None) -> lib.Weight:
        """[Cout,Cin,3,3] -> weights of the data-gradient conv (channels transposed, taps flipped)."""
        wt = weight.detach().to(device=self.device, dtype=torch.float32).permute(1, 0, 2, 3).flip(2, 3).contiguous()
        cout = weight.shape[0]
        cp = _ceil(cout if cout_pad is None else cout_pad, 16)
        return self._W(pack_conv3x3(wt, cin_pad=cp), weight.shape[1], 9 * cp)

    def _pack_all(self, lanes: int = 0):
        """forward + data-gradient weight images; the ~300 image packs are deferred and issued as a handful of launches.
        `lanes` > 0 (graph capture only): fork that many streams off the current one, deal the per-module re-layouts to
        them (Engine._pack_lane) and join before the image packs."""
        lib.PACK_QUEUE = []
        main = torch.cuda.current_stream()
        if lanes > 0:
            if len(self.__dict__.setdefault("_lane_streams", [])) < lanes:
                self._lane_streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]
            self._pack_streams, self._pack_main = self._lane_streams[:lanes], main
            for s in self._pack_streams:
                s.wait_stream(main)
        try:
            self.packed = self._pack()
            self._pack_train()
        finally:
            if lanes > 0:
                torch.cuda.set_stream(main)
                for s in self._pack_streams:
                    main.wait_stream(s)
                self._pack_streams = self._pack_main = None
            lib.flush_packs()
        self._train_packed_version = self.packed

    def _ensure_packed(self):
        v = self._param_versions()
        if self.packed is None or v != self._versions or self._train_packed_version is not self.packed:
            self._pack_all()
            self._versions = v
            self._graphs.clear()

    @torch.no_grad()
    def _pack_train(self):
        net, cfg, P = self.net, self.cfg, self.packed
        f32 = lambda t: t.detach().to(device=self.device, dtype=torch.float32)  # noqa: E731
        lane = 0   # the same module -> lane dealing as Engine._pack: a lane continues from its own results
        for st in cfg.stages():
            for blk, d in zip(getattr(net, st.name).blocks, P[st.name]):
                self._pack_lane(lane)
                lane += 1
                for k in ("qkv_w", "proj_w", "sqkv_w", "fc1_w", "fc2_w"):
                    d[k + ".d"] = self._dgrad_of(d[k])
                d["sdw.f"] = _flip_dw(d["sdw"])
                l = blk.local_spectral_attn
                # ll_w = [W_prompt ; W_down ; 0], rows padded to a multiple of 8 (the GEMM engine's K granularity when the
                # matrix is used transposed for dL/dm) — packed by Engine._pack (fold_weights = False)
                d["ll_w.d"] = self._dgrad_of(d["ll_w"])
                d["gate_raw"] = {
                    "param": f32(l.prompt_param).reshape(PROMPT_LEN, st.rank).contiguous(),
                    "q": f32(l.q.weight).contiguous(), "kv": f32(l.kv.weight).contiguous(),
                    "proj": f32(l.proj.weight).contiguous(), "proj_bias": f32(l.proj.bias).contiguous(),
                    "up": f32(l.linear_up.weight).contiguous(),
                }
                d["sout"] = f32(blk.gobal_spectral_attn.project_out.weight).reshape(st.dim, st.dim).contiguous()
        for name in ("prompt1", "prompt2"):
            self._pack_lane(lane)
            lane += 1
            m, d = getattr(net, name), P[name]
            for k in ("q_w", "kv_w", "pin_w", "pout_w"):
                d[k + ".d"] = self._dgrad_of(d[k])
            d["q_dw.f"], d["kv_dw.f"], d["ffn_dw.f"] = _flip_dw(d["q_dw"]), _flip_dw(d["kv_dw"]), _flip_dw(d["ffn_dw"])
            d["out"] = f32(m.cross_transformer.attn.project_out.weight).reshape(d["D"], d["D"]).contiguous()
            d["conv_last.d"] = self._conv_dgrad(m.conv_last.weight)
        for name in ("fusion1", "fusion2"):
            self._pack_lane(lane)
            lane += 1
            m, d = getattr(net, name), P[name]
            for k in ("qkv_w", "pin_w", "pout_w", "conv_w"):
                d[k + ".d"] = self._dgrad_of(d[k])
            d["dw.f"], d["ffn_dw.f"] = _flip_dw(d["dw"]), _flip_dw(d["ffn_dw"])
            d["out"] = f32(m.transformer.attn.project_out.weight).reshape(d["C"], d["C"]).contiguous()
        self._pack_lane(-1)
        P["reduce_chan_level2.d"] = self._dgrad_of(P["reduce_chan_level2"])
        P["cout_p"] = _ceil(cfg.out_channel, 16)
        P["output.d"] = self._conv_dgrad(net.output.weight, cout_pad=P["cout_p"])
        P["up2_1.d"] = self._conv_dgrad(net.up2_1.body[0].weight)
        P["up3_2.d"] = self._conv_dgrad(net.up3_2.body[0].weight)
        P["down2_3.d"] = self._conv_dgrad(net.down2_3.body[0].weight)
        P["down1_2.d"] = self._conv_dgrad(net.down1_2.body[0].weight)

    # -- small helpers ---------------------------------------------------------------------------------
    def _wgrad(self, dY: View, X: View, name: str, **kw):
        lib.wgrad(dY, X, self.g[name], self.prec, **kw)

    def _conv_wgrad(self, dY: View, X: View, name: str, cout: int, cin: int, H: int, W: int):
        lib.wgrad(dY, X, self.g[name], self.prec, taps=9, H=H, W=W, so=9 * cin, si=9, st=1, map_a=cout, i_valid=cin)

    def _scaled(self, tag: str, d: View, s: Optional[torch.Tensor], rpb: int) -> View:
        """DropPath: gradient of `s_b * branch` w.r.t. branch."""
        if s is None:
            return d
        out = self.ws.mat(tag, d.rows, d.cols)
        lib.axpby(d, out, 1.0, 0.0, row_scale=s, rows_per_batch=rpb)
        return out

    # -- spectral attention (shared by PGSSTB, PromptFusion, TVSP) --------------------------------------
    def _spectral_fwd(self, tag: str, q: View, q_shared: bool, k: View, k_shared: bool, temp, out_t, B, HW, heads, c):
        """Gram -> softmax -> fold.  Keeps the reduced statistics and the fp32 folded matrix for the backward."""
        ws = self.ws
        C = heads * c
        nfl, nch = lib.gram_partial_floats(B, heads, c, HW)
        partial = ws.flat(tag + ".partial", nfl)
        lib.gram_partial(q, q_shared, k, k_shared, partial, B, HW, heads, c)
        return self._spectral_finish_train(tag, partial, nch, temp, out_t, B, heads, c)

    def _dwgram_fwd(self, tag: str, t3: View, w_dw, temp, out_t, B, H, W, C, heads):
        """fused depthwise conv + Gram (forward kernel of the inference path): only v is written; q and k are recomputed by
        one depthwise conv in the backward.  Returns (v, spectral state)."""
        ws = self.ws
        c = C // heads
        v = ws.mat(tag + ".v", B * H * W, C)
        nfl, nch = lib.dwgram_partial_floats(B, heads, c, H, W)
        partial = ws.flat(tag + ".partial", nfl)
        lib.dwgram(t3, w_dw, v, partial, B, H, W, C, heads, self.prec)
        return v, self._spectral_finish_train(tag, partial, nch, temp, out_t, B, heads, c)

    def _spectral_finish_train(self, tag: str, partial, nch: int, temp, out_t, B, heads, c):
        ws = self.ws
        C = heads * c
        gsum = ws.flat(tag + ".gsum", B * heads * (c * c + 2 * c))
        ldm = _ldb(C)
        Mt = ws.flat(tag + ".Mt", B * _ceil(C, 16) * ldm)[: B * _ceil(C, 16) * ldm].view(B, _ceil(C, 16), ldm)
        nb = lib.bimg_bytes(C, C)
        img = ws.raw(tag + ".img", B * nb)[: B * nb].view(B, nb)
        lib.spectral_finish(partial, nch, gsum, temp, out_t, Mt, img, B, heads, c)
        return {"gsum": gsum if nch > 1 else partial, "Mt": Mt, "w": lib.Weight(None, img, C, C)}

    def _spectral_bwd(self, tag: str, S: dict, du: View, qk: View, v: View, v_shared_rows: int, out_w: torch.Tensor,
                      temp: torch.Tensor, g_out: torch.Tensor, g_temp: torch.Tensor, B: int, HW: int, heads: int, c: int,
                      dqk: View, dv: View):
        """Given dU = dL/d(project_out(A v)) [B*HW, C]:  dqk <- [dq | dk] (per sample), dv <- dU (Wout blockdiag(A))."""
        ws = self.ws
        C = heads * c
        Pm = ws.flat("b.P", B * C * C)[: B * C * C]
        Pm.zero_()
        lib.wgrad(du, v, Pm, self.prec, so=C, rows_per_batch=HW, dw_batch_stride=C * C, x_row_mod=v_shared_rows)
        ldwb = _ldb(2 * C)
        Wb = ws.flat("b.Wb", B * 2 * C * ldwb)[: B * 2 * C * ldwb].view(B, 2 * C, ldwb)
        lib.spectral_bwd(Pm, out_w, S["gsum"], temp, Wb, g_out, g_temp, B, heads, c)
        nb = lib.bimg_bytes(2 * C, 2 * C)
        wimg = ws.raw("b.Wbimg", B * nb)[: B * nb].view(B, nb)
        lib.pack_bimg(Wb, 2 * C, 2 * C, transposed=True, img=wimg)
        lib.gemm(qk, lib.Weight(None, wimg, 2 * C, 2 * C), dqk, 2 * C, precision=self.prec, rows_per_batch=HW)
        nb2 = lib.bimg_bytes(C, C)
        mimg = ws.raw("b.Mdimg", B * nb2)[: B * nb2].view(B, nb2)
        lib.pack_bimg(S["Mt"], C, C, transposed=False, img=mimg)
        lib.gemm(du, lib.Weight(None, mimg, C, C), dv, C, precision=self.prec, rows_per_batch=HW)

    # -- PGSSTB ---------------------------------------------------------------------------------------------
    def _block_fwd(self, pre: str, w: dict, st: Stage, shift: int, x: View, out: View, res2: Optional[View], B, H, W,
                   scales) -> dict:
        ws = self.ws
        C, heads = st.dim, st.heads
        N = B * H * W
        B_ = N // 64
        s1 = None if scales is None else scales[0]
        s2 = None if scales is None else scales[1]
        S = {"x": x, "shift": shift, "s1": s1, "s2": s2}
        y1 = S["y1"] = ws.mat(pre + "y1", N, C)
        S["st1"] = ws.flat(pre + "st1", 2 * N)
        lib.layernorm_fwd(x, w["ln1"], y1, S["st1"])
        qkv = S["qkv"] = ws.mat(pre + "qkv", N, 3 * C)
        self._gemm(y1, w["qkv_w"], qkv, 3 * C, bias=w["qkv_b"])
        core = S["core"] = ws.mat(pre + "core", N, C)
        wmean = ws.flat("wmean", B_ * C)
        lib.window_attn(qkv, w["rpb"], core, wmean, B, H, W, C, heads, shift, precision=self.prec)
        sa = S["sa"] = ws.mat(pre + "sa", N, C)
        self._gemm(core, w["proj_w"], sa, C, bias=w["proj_b"])
        gate = S["gate"] = ws.flat(pre + "gate", B_ * C)
        S["msa"] = ws.flat(pre + "msa", B_ * C)
        S["LL"] = ws.mat(pre + "LL", B_, _ceil(w["ll_w"].n, 16))
        self._gate_from_sa(w, st, sa, gate, B, H, W, shift, S["msa"], S["LL"])
        t3 = S["t3"] = ws.mat(pre + "t3", N, 3 * C)
        self._gemm(sa, w["sqkv_w"], t3, 3 * C)
        if lib.dwgram_supported(C, C // heads):
            S["dw3"] = None  # q, k recomputed in the backward
            v, S["spec"] = self._dwgram_fwd(pre + "spec", t3, w["sdw"], w["temp"], w["sout_t"], B, H, W, C, heads)
        else:
            dw3 = S["dw3"] = ws.mat(pre + "dw3", N, 3 * C)
            lib.dwconv3x3(t3, w["sdw"], dw3, B, H, W, 3 * C)
            S["spec"] = self._spectral_fwd(pre + "spec", dw3.cols_slice(0, C), False, dw3.cols_slice(C, 2 * C), False, w["temp"],
                                           w["sout_t"], B, H * W, heads, C // heads)
            v = dw3.cols_slice(2 * C, 3 * C)
        # mid = x + s1 * (sa * gate[win] + project_out(A v)): the gated part as one elementwise pass, the spectral part
        # through the TMA-drained residual epilogue (3.5 TB/s; the register-staged spectral epilogue runs at 1.8 TB/s)
        mid = S["mid"] = ws.mat(pre + "mid", N, C)
        u = ws.mat("u", N, C)
        lib.gate_apply_fwd(x, sa, gate, s1, u, B, H, W, C, shift)
        self._gemm(v, S["spec"]["w"], mid, C, epi=lib.EPI_RESIDUAL, res1=u, rows_per_batch=H * W, row_scale=s1)
        if lib.mlp_supported(C, w["hid_pad"]):
            lib.mlp(mid, w["ln2"], w["fc1_w"], w["fc1_b"], w["fc2_w"], w["fc2_b"], out, w["hid_pad"], self.prec,
                    res2=res2, row_scale=s2, rows_per_batch=H * W)
        else:
            hidden = ws.mat("hidden", N, w["hid_pad"])
            self._gemm(mid, w["fc1_w"], hidden, 2 * w["hid_pad"], ln=w["ln2"], bias=w["fc1_b"], epi=lib.EPI_GLU)
            self._gemm(hidden, w["fc2_w"], out, C, bias=w["fc2_b"], epi=lib.EPI_RESIDUAL, res1=mid, res2=res2,
                       rows_per_batch=H * W, row_scale=s2)
        return S

    def _block_bwd(self, pn: str, w: dict, st: Stage, S: dict, d_out: View, dx: View, B, H, W):
        """pn: parameter-name prefix '<stage>.blocks.<i>.'.  d_out -> dx (gradient w.r.t. the block input)."""
        ws, g = self.ws, self.g
        C, heads, r = st.dim, st.heads, st.rank
        hid, hp = self.cfg.hidden(C), w["hid_pad"]
        N = B * H * W
        B_ = N // 64
        HW = H * W
        shift, x, mid, sa = S["shift"], S["x"], S["mid"], S["sa"]
        # ---- x_out = mid + s2 * (fc2(value*gelu(gate)) + b2),  [value|gate] = fc1(LN2(mid)) + b1   (:719, :76-82)
        dm = self._scaled("b.dm", d_out, S["s2"], HW)
        y2 = ws.mat("b.y2", N, C)
        st2 = ws.flat("b.st2", 2 * N)
        lib.layernorm_fwd(mid, w["ln2"], y2, st2)
        h = ws.mat("b.h", N, 2 * hp)
        self._gemm(y2, w["fc1_w"], h, 2 * hp, bias=w["fc1_b"])
        dhid = ws.mat("b.dhid", N, hp)
        self._gemm(dm, w["fc2_w.d"], dhid, hp)
        lib.glu_bwd(h, dhid, hp)                      # h <- d(fc1 out), dhid <- hidden
        # bias gradients ride on the weight-gradient launches (column sums of the dY tiles, exact fp32)
        self._wgrad(dm, dhid, pn + "mlp.fc2.weight", i_valid=hid, dbias=g[pn + "mlp.fc2.bias"])
        self._wgrad(h, y2, pn + "mlp.fc1.weight", map_mode=MAP_INTERLEAVE, map_a=hid, dbias=g[pn + "mlp.fc1.bias"])
        gln = ws.mat("b.gln", N, C)
        self._gemm(h, w["fc1_w.d"], gln, C)
        d_mid = ws.mat("b.dmid", N, C)
        lib.layernorm_bwd(mid, st2, w["ln2"][0], gln, d_out, d_mid, g[pn + "norm2.weight"], g[pn + "norm2.bias"])
        # ---- mid = x + s1 * (sa * gate[win] + project_out(A v))                                     (:715-718)
        du = self._scaled("b.du", d_mid, S["s1"], HW)
        dw3, t3 = S["dw3"], S["t3"]
        if dw3 is None:  # forward ran the fused dwconv+Gram kernel: q, k, v = dwconv3x3(t3) again
            dw3 = ws.mat("b.dw3", N, 3 * C)
            lib.dwconv3x3(t3, w["sdw"], dw3, B, H, W, 3 * C)
        ddw3 = ws.mat("b.ddw3", N, 3 * C)
        sp = pn + "gobal_spectral_attn."
        self._spectral_bwd("b.spec", S["spec"], du, dw3.cols_slice(0, 2 * C), dw3.cols_slice(2 * C, 3 * C), 0, w["sout"],
                           w["temp"], g[sp + "project_out.weight"], g[sp + "temperature"], B, HW, heads, C // heads,
                           ddw3.cols_slice(0, 2 * C), ddw3.cols_slice(2 * C, 3 * C))
        dt3 = ws.mat("b.dt3", N, 3 * C)
        lib.dwconv3x3(ddw3, w["sdw.f"], dt3, B, H, W, 3 * C)
        lib.dwconv3x3_wgrad(t3, ddw3, g[sp + "qkv_dwconv.weight"], B, H, W, 3 * C)
        self._wgrad(dt3, sa, sp + "qkv.weight")
        # ---- local spectral gate (:135-153)
        dg = ws.flat("b.dg", B_ * C)
        lib.window_reduce(du, sa, dg, B, H, W, C, shift, 1.0)
        msa, LL = S["msa"], S["LL"]     # window mean of sa and [W_prompt m | W_down m], kept by the forward
        msa_v, dg_v = View(msa.data_ptr(), C, B_, C, msa), View(dg.data_ptr(), C, B_, C, dg)
        nl = w["ll_w"].n                # record columns [128+r, nl) hit zero weight rows
        ldr = lib.local_gate_bwd_record_ld(r)
        rec = ws.mat("b.rec", B_, ldr)
        lib.local_gate_bwd(LL, dg, w["gate_raw"], rec, B_, C, r)
        dmean = ws.flat("b.dmean", B_ * C)
        self._gemm(rec.cols_slice(0, nl), w["ll_w.d"], View(dmean.data_ptr(), C, B_, C, dmean), C)
        lp = pn + "local_spectral_attn."
        o_w, o_dsp, o_dq, o_sp, o_dkv, o_low, o_du, o_o, o_u = (128 + r, 256 + r, 256 + 2 * r, 256 + 3 * r, 256 + 4 * r,
                                                                 256 + 6 * r, 256 + 7 * r, 256 + 8 * r, 256 + 9 * r)
        # the seven r-sized weight gradients (+ proj.bias) are token-contractions of record columns: one launch
        R = rec.cols_slice
        lib.wgrad_multi([
            (R(0, 128), msa_v, g[lp + "linear_prompt.weight"], {}),
            (R(128, 128 + r), msa_v, g[lp + "linear_down.weight"], {}),
            (R(o_w, o_w + 128), R(o_dsp, o_dsp + r), g[lp + "prompt_param"], {}),
            (R(o_dq, o_dq + r), R(o_sp, o_sp + r), g[lp + "q.weight"], {}),
            (R(o_dkv, o_dkv + 2 * r), R(o_low, o_low + r), g[lp + "kv.weight"], {}),
            (R(o_du, o_du + r), R(o_o, o_o + r), g[lp + "proj.weight"], {"dbias": g[lp + "proj.bias"]}),
            (dg_v, R(o_u, o_u + r), g[lp + "linear_up.weight"], {}),
        ], self.prec)
        dsa0 = ws.mat("b.dsa0", N, C)
        lib.gate_apply_bwd(du, S["gate"], dmean, dsa0, B, H, W, C, shift)
        dsa = ws.mat("b.dsa", N, C)
        self._gemm(dt3, w["sqkv_w.d"], dsa, C, epi=lib.EPI_RESIDUAL, res1=dsa0)
        # ---- sa = proj(core) + b;  core = window attention(qkv);  qkv = LN1(x) Wqkv + b              (:195-216, :667)
        ap = pn + "attn."
        self._wgrad(dsa, S["core"], ap + "proj.weight", dbias=g[ap + "proj.bias"])
        dcore = ws.mat("b.dcore", N, C)
        self._gemm(dsa, w["proj_w.d"], dcore, C)
        dqkv = ws.mat("b.dqkv", N, 3 * C)
        groups = lib.window_attn_bwd_groups(B, H, W, heads)
        partial = ws.flat("b.wabp", groups * heads * 4096)
        lib.window_attn_bwd(S["qkv"], w["rpb"], dcore, dqkv, partial, groups, B, H, W, C, heads, shift, precision=self.prec)
        dbias = ws.flat("b.dbias", heads * 4096)[: heads * 4096]
        dbias.zero_()
        lib.colsum(View(partial.data_ptr(), heads * 4096, groups, heads * 4096, partial), dbias)
        lib.rpb_table_bwd(dbias, g[ap + "relative_position_bias_table"], heads)
        self._wgrad(dqkv, S["y1"], ap + "qkv.weight", dbias=g[ap + "qkv.bias"])
        gln1 = ws.mat("b.gln", N, C)
        self._gemm(dqkv, w["qkv_w.d"], gln1, C)
        lib.layernorm_bwd(x, S["st1"], w["ln1"][0], gln1, d_mid, dx, g[pn + "norm1.weight"], g[pn + "norm1.bias"])

    def _stage_fwd(self, name: str, x_in: View, out: View, B, H, W, keep) -> list:
        st = {s.name: s for s in self.cfg.stages()}[name]
        saved, x = [], x_in
        N = B * H * W
        for i, w in enumerate(self.packed[name]):
            last = i == len(self.packed[name]) - 1
            pre = f"{name}.{i}."
            dst = out if last else self.ws.mat(pre + "out", N, st.dim)
            scales = None if keep is None else keep.get((name, i))
            saved.append(self._block_fwd(pre, w, st, SHIFT if i % 2 else 0, x, dst, x_in if last else None, B, H, W, scales))
            x = dst
        return saved

    def _stage_bwd(self, name: str, saved: list, d_out: View, d_in: View, B, H, W):
        """BaseBlock: out = blocks(x) + x (net/MP_HSIR.py:760).  d_in <- dL/dx."""
        st = {s.name: s for s in self.cfg.stages()}[name]
        N = B * H * W
        ping = [self.ws.mat("b.gping", N, st.dim), self.ws.mat("b.gpong", N, st.dim)]
        d = d_out
        n = len(saved)
        for i in reversed(range(n)):
            dst = d_in if i == 0 else ping[i % 2]
            self._block_bwd(f"{name}.blocks.{i}.", self.packed[name][i], st, saved[i], d, dst, B, H, W)
            d = dst
        lib.axpby(d_out, d_in, 1.0, 1.0)

    # -- GDFN (FeedForward / FFN, :386-391, :260-265) ---------------------------------------------------------
    def _gdfn_fwd(self, tag: str, w: dict, x: View, out: View, ln, B, H, W, D) -> dict:
        ws = self.ws
        N, hp = B * H * W, w["hid_pad"]
        S = {"x": x}
        y = S["y"] = ws.mat(tag + ".y", N, D)
        S["st"] = ws.flat(tag + ".st", 2 * N)
        lib.layernorm_fwd(x, ln, y, S["st"])
        hin = S["hin"] = ws.mat(tag + ".hin", N, 2 * hp)
        self._gemm(y, w["pin_w"], hin, 2 * hp)
        dwo = S["dwo"] = ws.mat(tag + ".dwo", N, 2 * hp)
        lib.dwconv3x3(hin, w["ffn_dw"], dwo, B, H, W, 2 * hp)
        hg = S["hg"] = ws.mat(tag + ".hg", N, hp)
        lib.gdfn_gate_fwd(dwo, hg, hp)
        self._gemm(hg, w["pout_w"], out, D, epi=lib.EPI_RESIDUAL, res1=x)
        return S

    def _gdfn_bwd(self, pn: str, ln_name: str, w: dict, S: dict, ln, d_out: View, dx: View, B, H, W, D):
        """out = x + pout(gelu(a)*b): d_out -> dx.  pn: '<module>.ffn.', ln_name: '<module>.norm2.body.'"""
        ws, g = self.ws, self.g
        N, hp, hid = B * H * W, w["hid_pad"], self.cfg.hidden(D)
        self._wgrad(d_out, S["hg"], pn + "project_out.weight", i_valid=hid)
        dhg = ws.mat("b.f.dhg", N, hp)
        self._gemm(d_out, w["pout_w.d"], dhg, hp)
        ddwo = ws.mat("b.f.ddwo", N, 2 * hp)
        lib.gdfn_gate_bwd(S["dwo"], dhg, ddwo, hp)
        dhin = ws.mat("b.f.dhin", N, 2 * hp)
        lib.dwconv3x3(ddwo, w["ffn_dw.f"], dhin, B, H, W, 2 * hp)
        lib.dwconv3x3_wgrad(S["hin"], ddwo, g[pn + "dwconv.weight"], B, H, W, 2 * hp, MAP_HALVES, hid, hp)
        self._wgrad(dhin, S["y"], pn + "project_in.weight", map_mode=MAP_HALVES, map_a=hid, map_b=hp)
        gl = ws.mat("b.f.gl", N, D)
        self._gemm(dhin, w["pin_w.d"], gl, D)
        lib.layernorm_bwd(S["x"], S["st"], ln[0], gl, d_out, dx, g[ln_name + "weight"], g[ln_name + "bias"])

    # -- PromptFusion (:594-599 -> TransformerBlock :475-479) --------------------------------------------------
    def _fusion_fwd(self, name: str, xcat: View, out: View, B, H, W) -> dict:
        w, ws = self.packed[name], self.ws
        C2, heads = w["C"], w["heads"]
        N = B * H * W
        S = {"xcat": xcat}
        y = S["y"] = ws.mat(name + ".y", N, C2)
        S["st"] = ws.flat(name + ".st", 2 * N)
        lib.layernorm_fwd(xcat, w["ln1"], y, S["st"])
        t3 = S["t3"] = ws.mat(name + ".t3", N, 3 * C2)
        self._gemm(y, w["qkv_w"], t3, 3 * C2)
        if lib.dwgram_supported(C2, C2 // heads):
            S["dw3"] = None
            v, S["spec"] = self._dwgram_fwd(name + ".spec", t3, w["dw"], w["temp"], w["out_t"], B, H, W, C2, heads)
        else:
            dw3 = S["dw3"] = ws.mat(name + ".dw3", N, 3 * C2)
            lib.dwconv3x3(t3, w["dw"], dw3, B, H, W, 3 * C2)
            S["spec"] = self._spectral_fwd(name + ".spec", dw3.cols_slice(0, C2), False, dw3.cols_slice(C2, 2 * C2), False,
                                           w["temp"], w["out_t"], B, H * W, heads, C2 // heads)
            v = dw3.cols_slice(2 * C2, 3 * C2)
        y1 = S["y1"] = ws.mat(name + ".y1", N, C2)
        self._gemm(v, S["spec"]["w"], y1, C2, epi=lib.EPI_RESIDUAL, res1=xcat, rows_per_batch=H * W)
        y2 = S["y2"] = ws.mat(name + ".y2", N, C2)
        S["ffn"] = self._gdfn_fwd(name + ".ffn", w, y1, y2, w["ln2"], B, H, W, C2)
        self._gemm(y2, w["conv_w"], out, out.cols)
        return S

    def _fusion_bwd(self, name: str, S: dict, d_out: View, d_xcat: View, B, H, W):
        w, ws, g = self.packed[name], self.ws, self.g
        C2, heads = w["C"], w["heads"]
        N, HW = B * H * W, H * W
        tb = name + ".transformer."
        self._wgrad(d_out, S["y2"], name + ".conv.weight")
        dy2 = ws.mat("b.u.dy2", N, C2)
        self._gemm(d_out, w["conv_w.d"], dy2, C2)
        dy1 = ws.mat("b.u.dy1", N, C2)
        self._gdfn_bwd(tb + "ffn.", tb + "norm2.body.", w, S["ffn"], w["ln2"], dy2, dy1, B, H, W, C2)
        dw3, t3 = S["dw3"], S["t3"]
        if dw3 is None:
            dw3 = ws.mat("b.u.dw3", N, 3 * C2)
            lib.dwconv3x3(t3, w["dw"], dw3, B, H, W, 3 * C2)
        ddw3 = ws.mat("b.u.ddw3", N, 3 * C2)
        self._spectral_bwd("b.spec", S["spec"], dy1, dw3.cols_slice(0, 2 * C2), dw3.cols_slice(2 * C2, 3 * C2), 0, w["out"],
                           w["temp"], g[tb + "attn.project_out.weight"], g[tb + "attn.temperature"], B, HW, heads, C2 // heads,
                           ddw3.cols_slice(0, 2 * C2), ddw3.cols_slice(2 * C2, 3 * C2))
        dt3 = ws.mat("b.u.dt3", N, 3 * C2)
        lib.dwconv3x3(ddw3, w["dw.f"], dt3, B, H, W, 3 * C2)
        lib.dwconv3x3_wgrad(t3, ddw3, g[tb + "attn.qkv_dwconv.weight"], B, H, W, 3 * C2)
        self._wgrad(dt3, S["y"], tb + "attn.qkv.weight")
        gl = ws.mat("b.u.gl", N, C2)
        self._gemm(dt3, w["qkv_w.d"], gl, C2)
        lib.layernorm_bwd(S["xcat"], S["st"], w["ln1"][0], gl, dy1, d_xcat, g[tb + "norm1.body.weight"], g[tb + "norm1.body.bias"])

    # -- TVSP (:572-583) ---------------------------------------------------------------------------------------
    def _tvsp_fwd(self, name: str, clip_b, weights, B, Hs, Ws, out: View) -> dict:
        w, ws = self.packed[name], self.ws
        D, ps = w["D"], w["ps"]
        T, n = self.cfg.task_classes, w["ps"] * w["ps"]
        S = {}
        Q = S["Q"] = ws.mat(name + ".Q", B * n, D)
        lib.tvsp_query(clip_b, weights, w["learnable"], Q, B, T, D, ps)
        yq = S["yq"] = ws.mat(name + ".yq", B * n, D)
        S["stq"] = ws.flat(name + ".stq", 2 * B * n)
        lib.layernorm_fwd(Q, w["ln11"], yq, S["stq"])
        q1 = S["q1"] = ws.mat(name + ".q1", B * n, D)
        self._gemm(yq, w["q_w"], q1, D)
        qd = S["qd"] = ws.mat(name + ".qd", B * n, D)
        lib.dwconv3x3(q1, w["q_dw"], qd, B, ps, ps, D)
        vis = S["vis"] = View.of(w["visual"])
        yv = S["yv"] = ws.mat(name + ".yv", n, D)
        S["stv"] = ws.flat(name + ".stv", 2 * n)
        lib.layernorm_fwd(vis, w["ln12"], yv, S["stv"])
        kv1 = S["kv1"] = ws.mat(name + ".kv1", n, 2 * D)
        self._gemm(yv, w["kv_w"], kv1, 2 * D)
        kvd = S["kvd"] = ws.mat(name + ".kvd", n, 2 * D)
        lib.dwconv3x3(kv1, w["kv_dw"], kvd, 1, ps, ps, 2 * D)
        S["spec"] = self._spectral_fwd(name + ".spec", qd, False, kvd.cols_slice(0, D), True, w["temp"], w["out_t"], B, n, 2, D // 2)
        xa = S["xa"] = ws.mat(name + ".xa", B * n, D)
        self._gemm(kvd.cols_slice(D, 2 * D), S["spec"]["w"], xa, D, epi=lib.EPI_RESIDUAL, res1=Q, rows_per_batch=n, a_row_mod=n,
                   M=B * n)
        pr = S["pr"] = ws.mat(name + ".pr", B * n, D)
        S["ffn"] = self._gdfn_fwd(name + ".ffn", w, xa, pr, w["ln2"], B, ps, ps, D)
        if (Hs, Ws) != (ps, ps):
            prr = ws.mat(name + ".prr", B * Hs * Ws, D)
            lib.bilinear(pr, prr, B, ps, ps, Hs, Ws, D)
            pr = prr
        S["conv_in"] = pr
        self._conv(pr, w["conv_last"], out.ptr, out.ld, B, Hs, Ws, D, D, lib.CONV_TOKENS)
        S["clip_b"], S["weights"] = clip_b, weights
        return S

    def _tvsp_bwd(self, name: str, S: dict, d_out: View, B, Hs, Ws):
        w, ws, g = self.packed[name], self.ws, self.g
        D, ps = w["D"], w["ps"]
        T, n = self.cfg.task_classes, w["ps"] * w["ps"]
        ct = name + ".cross_transformer."
        self._conv_wgrad(d_out, S["conv_in"], name + ".conv_last.weight", D, D, Hs, Ws)
        dprr = ws.mat("b.t.dprr", B * Hs * Ws, D)
        self._conv(d_out, w["conv_last.d"], dprr.ptr, dprr.ld, B, Hs, Ws, D, D, lib.CONV_TOKENS)
        if (Hs, Ws) != (ps, ps):
            dpr = ws.mat("b.t.dpr", B * n, D)
            dpr.keep.zero_()
            lib.bilinear_bwd(dprr, dpr, B, ps, ps, Hs, Ws, D)
        else:
            dpr = dprr
        dxa = ws.mat("b.t.dxa", B * n, D)
        self._gdfn_bwd(ct + "ffn.", ct + "norm2.body.", w, S["ffn"], w["ln2"], dpr, dxa, B, ps, ps, D)
        # xa = Q + project_out(A_b v), q from the text query (per sample), k / v from the visual prompt (shared)
        kvd, qd = S["kvd"], S["qd"]
        qk = ws.mat("b.t.qk", B * n, 2 * D)
        lib.axpby(qd, qk.cols_slice(0, D))
        lib.axpby(kvd.cols_slice(0, D), qk.cols_slice(D, 2 * D), x_row_mod=n)
        dqk = ws.mat("b.t.dqk", B * n, 2 * D)
        dvb = ws.mat("b.t.dvb", B * n, D)
        self._spectral_bwd("b.spec", S["spec"], dxa, qk, kvd.cols_slice(D, 2 * D), n, w["out"], w["temp"],
                           g[ct + "attn.project_out.weight"], g[ct + "attn.temperature"], B, n, 2, D // 2, dqk, dvb)
        dkvd = ws.mat("b.t.dkvd", n, 2 * D)
        lib.batch_sum(dqk.cols_slice(D, 2 * D), dkvd.cols_slice(0, D), B)
        lib.batch_sum(dvb, dkvd.cols_slice(D, 2 * D), B)
        dkv1 = ws.mat("b.t.dkv1", n, 2 * D)
        lib.dwconv3x3(dkvd, w["kv_dw.f"], dkv1, 1, ps, ps, 2 * D)
        lib.dwconv3x3_wgrad(S["kv1"], dkvd, g[ct + "attn.kv_dwconv.weight"], 1, ps, ps, 2 * D)
        self._wgrad(dkv1, S["yv"], ct + "attn.kv.weight")
        glv = ws.mat("b.t.glv", n, D)
        self._gemm(dkv1, w["kv_w.d"], glv, D)
        dvis = ws.mat("b.t.dvis", n, D)
        lib.layernorm_bwd(S["vis"], S["stv"], w["ln12"][0], glv, None, dvis, g[ct + "norm12.body.weight"], g[ct + "norm12.body.bias"])
        lib.tokens_to_nchw(dvis, g[name + ".visual_prompt"], 1, D, n)
        dqd = dqk.cols_slice(0, D)
        dq1 = ws.mat("b.t.dq1", B * n, D)
        lib.dwconv3x3(dqd, w["q_dw.f"], dq1, B, ps, ps, D)
        lib.dwconv3x3_wgrad(S["q1"], dqd, g[ct + "attn.q_dwconv.weight"], B, ps, ps, D)
        self._wgrad(dq1, S["yq"], ct + "attn.q.weight")
        glq = ws.mat("b.t.glq", B * n, D)
        self._gemm(dq1, w["q_w.d"], glq, D)
        dQ = ws.mat("b.t.dQ", B * n, D)
        lib.layernorm_bwd(S["Q"], S["stq"], w["ln11"][0], glq, dxa, dQ, g[ct + "norm11.body.weight"], g[ct + "norm11.body.bias"])
        lib.tvsp_query_bwd(dQ, S["clip_b"], S["weights"], g[name + ".text_prompt_learnable"], B, T, D, ps)

    # -- whole network ------------------------------------------------------------------------------------------
    def forward_train(self, x: torch.Tensor, weights: torch.Tensor, out: torch.Tensor, keep=None) -> dict:
        """MP_HSIR_Net.forward (net/MP_HSIR.py:810-844) keeping what the backward needs.  keep: {(stage, i): [2,B]}
        DropPath multipliers (mask/keep_prob, :718-719) or None."""
        cfg, P, ws = self.cfg, self.packed, self.ws
        self._tvsp_valid.clear()   # the prompt columns of the fusion buffers are rewritten below (inference-time cache)
        B, _, H, W = x.shape
        d = cfg.dim
        N1, N2, N3 = B * H * W, B * H * W // 4, B * H * W // 16
        H2, W2, H3, W3 = H // 2, W // 2, H // 4, W // 4
        T = cfg.task_classes
        F = {"shape": (B, H, W)}
        clip_b = ws.flat("clip_b", B * 512)
        lib.text_prompt(weights, P["clip"], clip_b, B, T)
        tok = F["tok"] = ws.mat("tok_in", N1, P["cin_p"])
        lib.nchw_to_tokens(x, tok)
        x1 = ws.mat("x1", N1, d)
        self._conv(tok, P["patch_embed"], x1.ptr, x1.ld, B, H, W, P["cin_p"], d)
        fcat1 = F["fcat1"] = ws.mat("fcat1", N1, 2 * d)
        e1 = fcat1.cols_slice(0, d)
        F["enc1"] = self._stage_fwd("encoder_level1", x1, e1, B, H, W, keep)
        x2 = ws.mat("x2", N2, 2 * d)
        self._conv(e1, P["down1_2"], x2.ptr, x2.ld, B, H, W, d, d // 2, lib.CONV_UNSHUFFLE)
        fcat2 = F["fcat2"] = ws.mat("fcat2", N2, 4 * d)
        e2 = fcat2.cols_slice(0, 2 * d)
        F["enc2"] = self._stage_fwd("encoder_level2", x2, e2, B, H2, W2, keep)
        x3 = ws.mat("x3", N3, 4 * d)
        self._conv(e2, P["down2_3"], x3.ptr, x3.ld, B, H2, W2, 2 * d, d, lib.CONV_UNSHUFFLE)
        lat = F["lat"] = ws.mat("lat", N3, 4 * d)
        F["latent"] = self._stage_fwd("latent", x3, lat, B, H3, W3, keep)
        cat2 = F["cat2"] = ws.mat("cat2", N2, 4 * d)
        self._conv(lat, P["up3_2"], cat2.ptr, cat2.ld, B, H3, W3, 4 * d, 8 * d, lib.CONV_SHUFFLE)
        F["prompt2"] = self._tvsp_fwd("prompt2", clip_b, weights, B, H2, W2, fcat2.cols_slice(2 * d, 4 * d))
        F["fusion2"] = self._fusion_fwd("fusion2", fcat2, cat2.cols_slice(2 * d, 4 * d), B, H2, W2)
        d2in = ws.mat("d2in", N2, 2 * d)
        self._gemm(cat2, P["reduce_chan_level2"], d2in, 2 * d)
        d2 = F["d2"] = ws.mat("d2", N2, 2 * d)
        F["dec2"] = self._stage_fwd("decoder_level2", d2in, d2, B, H2, W2, keep)
        cat1 = ws.mat("cat1", N1, 2 * d)
        self._conv(d2, P["up2_1"], cat1.ptr, cat1.ld, B, H2, W2, 2 * d, 4 * d, lib.CONV_SHUFFLE)
        F["prompt1"] = self._tvsp_fwd("prompt1", clip_b, weights, B, H, W, fcat1.cols_slice(d, 2 * d))
        F["fusion1"] = self._fusion_fwd("fusion1", fcat1, cat1.cols_slice(d, 2 * d), B, H, W)
        dd1 = ws.mat("dd1", N1, 2 * d)
        F["dec1"] = self._stage_fwd("decoder_level1", cat1, dd1, B, H, W, keep)
        ref = F["ref"] = ws.mat("ref", N1, 2 * d)
        F["refine"] = self._stage_fwd("refinement", dd1, ref, B, H, W, keep)
        self._conv(ref, P["output"], out.data_ptr(), 0, B, H, W, 2 * d, cfg.out_channel, lib.CONV_NCHW_RES, R=x)
        return F

    def backward(self, F: dict, d_out: torch.Tensor):
        """d_out: dL/d(output) NCHW.  Accumulates every parameter gradient into the flat gradient buffer."""
        cfg, P, ws = self.cfg, self.packed, self.ws
        B, H, W = F["shape"]
        d = cfg.dim
        N1, N2, N3 = B * H * W, B * H * W // 4, B * H * W // 16
        H2, W2, H3, W3 = H // 2, W // 2, H // 4, W // 4
        co, cop = cfg.out_channel, P["cout_p"]
        dtok = ws.mat("b.dtok", N1, cop)
        lib.nchw_to_tokens(d_out, dtok)
        self._conv_wgrad(dtok, F["ref"], "output.weight", co, 2 * d, H, W)
        d_ref = ws.mat("b.n.dref", N1, 2 * d)
        self._conv(dtok, P["output.d"], d_ref.ptr, d_ref.ld, B, H, W, cop, 2 * d, lib.CONV_TOKENS)
        d_dd1 = ws.mat("b.n.ddd1", N1, 2 * d)
        self._stage_bwd("refinement", F["refine"], d_ref, d_dd1, B, H, W)
        d_cat1 = ws.mat("b.n.dcat1", N1, 2 * d)
        self._stage_bwd("decoder_level1", F["dec1"], d_dd1, d_cat1, B, H, W)
        # cat1 = [PixelShuffle(up2_1(d2)) | fusion1(e1, prompt1)]
        d_fcat1 = ws.mat("b.n.dfcat1", N1, 2 * d)
        self._fusion_bwd("fusion1", F["fusion1"], d_cat1.cols_slice(d, 2 * d), d_fcat1, B, H, W)
        self._tvsp_bwd("prompt1", F["prompt1"], d_fcat1.cols_slice(d, 2 * d), B, H, W)
        t21 = ws.mat("b.n.t21", N2, 4 * d)
        lib.pixel_unshuffle(d_cat1.cols_slice(0, d), t21, B, H, W, d)
        self._conv_wgrad(t21, F["d2"], "up2_1.body.0.weight", 4 * d, 2 * d, H2, W2)
        d_d2 = ws.mat("b.n.dd2", N2, 2 * d)
        self._conv(t21, P["up2_1.d"], d_d2.ptr, d_d2.ld, B, H2, W2, 4 * d, 2 * d, lib.CONV_TOKENS)
        d_d2in = ws.mat("b.n.dd2in", N2, 2 * d)
        self._stage_bwd("decoder_level2", F["dec2"], d_d2, d_d2in, B, H2, W2)
        self._wgrad(d_d2in, F["cat2"], "reduce_chan_level2.weight")
        d_cat2 = ws.mat("b.n.dcat2", N2, 4 * d)
        self._gemm(d_d2in, P["reduce_chan_level2.d"], d_cat2, 4 * d)
        d_fcat2 = ws.mat("b.n.dfcat2", N2, 4 * d)
        self._fusion_bwd("fusion2", F["fusion2"], d_cat2.cols_slice(2 * d, 4 * d), d_fcat2, B, H2, W2)
        self._tvsp_bwd("prompt2", F["prompt2"], d_fcat2.cols_slice(2 * d, 4 * d), B, H2, W2)
        t32 = ws.mat("b.n.t32", N3, 8 * d)
        lib.pixel_unshuffle(d_cat2.cols_slice(0, 2 * d), t32, B, H2, W2, 2 * d)
        self._conv_wgrad(t32, F["lat"], "up3_2.body.0.weight", 8 * d, 4 * d, H3, W3)
        d_lat = ws.mat("b.n.dlat", N3, 4 * d)
        self._conv(t32, P["up3_2.d"], d_lat.ptr, d_lat.ld, B, H3, W3, 8 * d, 4 * d, lib.CONV_TOKENS)
        d_x3 = ws.mat("b.n.dx3", N3, 4 * d)
        self._stage_bwd("latent", F["latent"], d_lat, d_x3, B, H3, W3)
        # x3 = PixelUnshuffle(down2_3(e2))
        e2, e1 = F["fcat2"].cols_slice(0, 2 * d), F["fcat1"].cols_slice(0, d)
        t23 = ws.mat("b.n.t23", N2, d)
        lib.pixel_shuffle(d_x3, t23, B, H3, W3, d)
        self._conv_wgrad(t23, e2, "down2_3.body.0.weight", d, 2 * d, H2, W2)
        d_e2 = ws.mat("b.n.de2", N2, 2 * d)
        self._conv(t23, P["down2_3.d"], d_e2.ptr, d_e2.ld, B, H2, W2, d, 2 * d, lib.CONV_TOKENS)
        lib.axpby(d_fcat2.cols_slice(0, 2 * d), d_e2, 1.0, 1.0)
        d_x2 = ws.mat("b.n.dx2", N2, 2 * d)
        self._stage_bwd("encoder_level2", F["enc2"], d_e2, d_x2, B, H2, W2)
        t12 = ws.mat("b.n.t12", N1, d // 2)
        lib.pixel_shuffle(d_x2, t12, B, H2, W2, d // 2)
        self._conv_wgrad(t12, e1, "down1_2.body.0.weight", d // 2, d, H, W)
        d_e1 = ws.mat("b.n.de1", N1, d)
        self._conv(t12, P["down1_2.d"], d_e1.ptr, d_e1.ld, B, H, W, d // 2, d, lib.CONV_TOKENS)
        lib.axpby(d_fcat1.cols_slice(0, d), d_e1, 1.0, 1.0)
        d_x1 = ws.mat("b.n.dx1", N1, d)
        self._stage_bwd("encoder_level1", F["enc1"], d_e1, d_x1, B, H, W)
        self._conv_wgrad(d_x1, F["tok"], "patch_embed.proj.weight", d, cfg.in_channel, H, W)

    # -- torch.autograd entry points (mp_hsir_b200/autograd.py) --------------------------------------------------
    autograd_serial = 0
    drop_path_generator: Optional[torch.Generator] = None   # tests: seeded DropPath draws
    _grads_attached = True   # _flatten_parameters points param.grad at views of flat_g (the train_step fast path)

    def release_param_grads(self):
        """hand ``param.grad`` back to torch: in autograd mode AccumulateGrad owns it (a grad that aliased the flat gradient
        buffer would be added to itself)"""
        if self._grads_attached:
            for p in self.net.parameters():
                p.grad = None
            self._grads_attached = False

    @torch.no_grad()
    def autograd_forward(self, inp: torch.Tensor, task_id: torch.Tensor):
        B, _, H, W = inp.shape
        if H % 32 or W % 32:
            raise ValueError(f"H and W must be multiples of 32, got {H}x{W}")
        self._ensure_packed()
        x = inp.detach().to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            out = torch.empty_like(x)
            keep = self.drop_path_scales(B, self.drop_path_generator) if self.net.training else None
            F = self.forward_train(x, self.task_weights(task_id), out, keep)
        TrainEngine.autograd_serial = self.autograd_serial = self.autograd_serial + 1
        return out.to(inp.dtype), F

    @torch.no_grad()
    def autograd_backward(self, F: dict, d_out: torch.Tensor) -> Dict[str, torch.Tensor]:
        """-> {parameter name: gradient tensor} (views of ONE fresh copy of the flat gradient buffer, so that they neither
        alias each other's future contents nor the next backward's)"""
        with torch.cuda.device(self.device):
            self.flat_g.zero_()
            self.backward(F, d_out.detach().to(torch.float32).contiguous())
            snap = self.flat_g.clone()
        params = dict(self.net.named_parameters())
        return {n: snap[o:o + params[n].numel()].view(params[n].shape) for n, o in self.param_offsets.items()}

    # -- public steps --------------------------------------------------------------------------------------------
    def drop_path_scales(self, B: int, generator: Optional[torch.Generator] = None) -> dict:
        """Per-sample DropPath multipliers mask/keep_prob for every block, one draw per call like timm's DropPath
        applied twice per block (net/MP_HSIR.py:718-719)."""
        slots = [(st.name, i) for st in self.cfg.stages() for i, rate in enumerate(st.dpr) if rate > 0.0]
        if not slots:
            return {}
        # every block's Bernoulli draws in four launches (one uniform draw for all of them) instead of four per block
        kp = self._dp_keep_prob   # built in __init__: a host-to-device copy must not happen inside a graph capture
        m = (torch.rand(len(slots), 2, B, device=self.device, generator=generator) < kp).to(torch.float32) / kp
        return {slot: m[j] for j, slot in enumerate(slots)}

    def _fwd_bwd(self, x, cl, weights, out, d_out, keep):
        """launch sequence of forward + clamp/L1 + backward on pre-allocated buffers (CUDA-graph capturable)."""
        self.flat_g.zero_()
        if keep == "sample":
            keep = self.drop_path_scales(x.shape[0])
        F = self.forward_train(x, weights, out, keep)
        self.loss_buf.zero_()
        lib.l1_clamp_loss(out, cl, d_out, self.loss_buf)
        self.backward(F, d_out)

    @torch.no_grad()
    def loss_and_grad(self, inp: torch.Tensor, clean: torch.Tensor, task_id: torch.Tensor, keep=None):
        """forward + clamp/L1 loss + backward.  Returns (restored, loss tensor[1]); gradients overwrite flat_g."""
        B, _, H, W = inp.shape
        if H % 32 or W % 32:
            raise ValueError(f"H and W must be multiples of 32, got {H}x{W}")
        self._ensure_packed()
        x = inp.detach().to(torch.float32).contiguous()
        cl = clean.detach().to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            out, d_out = torch.empty_like(x), torch.empty_like(x)
            self._fwd_bwd(x, cl, self.task_weights(task_id), out, d_out, keep)
        return out, self.loss_buf

    @torch.no_grad()
    def optimizer_step(self, grad_scale: float = 1.0):
        self.step_count += 1
        with torch.cuda.device(self.device):
            lib.adamw_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps,
                           self.weight_decay, self.step_count, grad_scale)
        # parameters changed in place behind torch's back: re-pack the weight images (captured steps stay valid)
        Engine.invalidate(self)

    def train_step(self, inp, clean, task_id, keep="sample", world_size: int = 1, all_reduce=None,
                   cuda_graph: bool = False) -> torch.Tensor:
        """One optimisation step (train.py:50-69): returns the loss tensor [1] (device).  With ``cuda_graph`` the launch
        sequence (~1.8k kernels) is captured once per input shape and replayed: weight re-packing, forward, loss,
        backward and AdamW become three graph launches around the (eager) gradient all-reduce."""
        if cuda_graph and lib.PROFILER is None:
            return self._train_step_graphed(inp, clean, task_id, keep, world_size, all_reduce)
        self.zero_grad()
        _, loss = self.loss_and_grad(inp, clean, task_id, keep)
        if all_reduce is not None:
            all_reduce(self.flat_g)  # DDP: sum over ranks, the mean is folded into grad_scale
        self.optimizer_step(grad_scale=1.0 / world_size)
        return loss.clone()   # loss_buf itself is overwritten by the next step

    # -- CUDA-graph replay of the step ---------------------------------------------------------------------------
    @torch.no_grad()
    def _train_step_graphed(self, inp, clean, task_id, keep, world_size, all_reduce):
        if keep not in (None, "sample"):
            raise ValueError("cuda_graph=True supports keep=None or 'sample' (masks are drawn inside the graph)")
        key = (tuple(inp.shape), keep, world_size)
        graphs = self.__dict__.setdefault("_train_graphs", {})
        ent = graphs.get(key)
        if ent is not None and ent != "warm" and ent[11] != self.ws.generation:
            # a named workspace buffer was re-allocated since the capture (a larger batch, or an inference forward at a
            # larger shape on this engine): every captured step holds raw pointers into the old blocks -> drop them all
            graphs.clear()
            ent = None
        if ent is None:
            # first call at this shape: eager step (sizes the workspace, sets kernel attributes); mark for capture
            graphs[key] = "warm"
            return self.train_step(inp, clean, task_id, keep, world_size, all_reduce, cuda_graph=False)
        with torch.cuda.device(self.device):
            if ent == "warm":
                ent = graphs[key] = self._capture_step(inp, clean, task_id, keep, world_size)
            g_pack, g_main, g_opt, sx, sc, sw, dyn = ent[:7]
            if self.packed is not ent[8]:
                # an eager step (or anything else that invalidated the packed weights) ran since the last replay: refresh the
                # graph's own weight images from the flat parameters and make them current again
                g_pack.replay()
                self.packed = self._train_packed_version = ent[8]
                self._versions = self._param_versions()
            sx.copy_(inp, non_blocking=True)
            sc.copy_(clean, non_blocking=True)
            sw.copy_(self.task_weights(task_id), non_blocking=True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if self.time_graphs else None
            if ev:
                ev[0].record()
            g_main.replay()
            if ev:
                ev[1].record()
            if all_reduce is not None:
                all_reduce(self.flat_g)
            self.step_count += 1
            dyn[0:1].fill_(self.lr)  # stream-ordered scalars: an lr schedule just changes self.lr
            dyn[3:4].fill_(float(self.step_count - 1))
            if ev:
                ev[2].record()
            g_opt.replay()           # bumps the device step counter, bias corrections, AdamW
            if ev:
                ev[3].record()
            g_pack.replay()          # weight images for the next forward
            if ev:
                ev[4].record()
                torch.cuda.synchronize()
                self.graph_ms = {"fwd_bwd": ev[0].elapsed_time(ev[1]), "all_reduce": ev[1].elapsed_time(ev[2]),
                                 "adamw": ev[2].elapsed_time(ev[3]), "repack": ev[3].elapsed_time(ev[4])}
            lib.LAUNCHES += ent[7]
        return self.loss_buf.clone()   # loss_buf itself is overwritten by the next step

    def _capture_step(self, inp, clean, task_id, keep, world_size):
        gen = self.ws.generation
        sx = torch.empty_like(inp, dtype=torch.float32).contiguous()
        sc = torch.empty_like(sx)
        sw = torch.empty_like(self.task_weights(task_id))
        out, d_out = torch.empty_like(sx), torch.empty_like(sx)
        dyn = torch.zeros(4, device=self.device, dtype=torch.float32)          # {lr, 1-b1^t, 1-b2^t, t}
        dyn[3] = float(self.step_count)
        sx.copy_(inp)
        sc.copy_(clean)
        sw.copy_(self.task_weights(task_id))
        torch.cuda.current_stream().synchronize()
        n0 = lib.LAUNCHES
        pool = torch.cuda.graph_pool_handle()
        # 1. weight (re-)packing from the flat parameter buffer: the packed tensors live in the graph's pool, so their
        #    addresses are what the forward/backward graph records
        g_pack = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_pack, pool=pool):
            self._pack_all(lanes=self.pack_lanes)
        self._versions = self._param_versions()
        n_pack = lib.LAUNCHES - n0
        g_pack.replay()
        # 2. zero_grad + DropPath masks + forward + loss + backward
        g_main = torch.cuda.CUDAGraph()
        n1 = lib.LAUNCHES
        with torch.cuda.graph(g_main, pool=pool):
            self._fwd_bwd(sx, sc, sw, out, d_out, keep)
        n_main = lib.LAUNCHES - n1
        if self.ws.generation != gen:
            raise RuntimeError("workspace grew during graph capture")
        # 3. AdamW reading {lr, bias corrections} from device memory
        g_opt = torch.cuda.CUDAGraph()
        n2 = lib.LAUNCHES
        with torch.cuda.graph(g_opt, pool=pool):
            dyn[3:4].add_(1.0)
            dyn[1:2].copy_(1.0 - torch.pow(self.betas[0], dyn[3:4]))
            dyn[2:3].copy_(1.0 - torch.pow(self.betas[1], dyn[3:4]))
            lib.adamw_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps,
                           self.weight_decay, 1, 1.0 / world_size, dyn=dyn)
        n_opt = lib.LAUNCHES - n2
        lib.LAUNCHES = n0  # capture records, it does not launch
        # `out` / `d_out` are written by every replay of g_main: they must live as long as the graph does
        return (g_pack, g_main, g_opt, sx, sc, sw, dyn, n_pack + n_main + n_opt, self.packed, out, d_out, gen)

    def invalidate(self):
        # a captured step owns its packed weights and refreshes them itself (g_pack); anything else (load_state_dict,
        # manual edits) drops the graphs
        super().invalidate()
        self.__dict__.pop("_train_graphs", None)
