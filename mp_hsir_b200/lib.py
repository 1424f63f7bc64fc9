"""ctypes binding of libmphsir.so (the C ABI in include/mphsir.h).

PyTorch is plumbing here: it owns device memory and streams; every compute call below goes
through the C ABI with raw device pointers.  There is NO fallback: if the library is missing or
a call fails, a RuntimeError carrying ``mphsir_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmphsir.so")

_lib: Optional[C.CDLL] = None
LAUNCHES = 0  # number of kernel-launching ABI calls made by this process (bench.py reports it)


class GemmParams(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int), ("a_row_mod", C.c_int),
        ("Bt", C.c_void_p), ("ldb", C.c_int),
        ("b_batch_stride", C.c_longlong), ("rows_per_batch", C.c_int),
        ("Y", C.c_void_p), ("ldy", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("bias", C.c_void_p), ("epi", C.c_int),
        ("res1", C.c_void_p), ("ldr1", C.c_int),
        ("res2", C.c_void_p), ("ldr2", C.c_int),
        ("gsrc", C.c_void_p), ("ldg", C.c_int),
        ("gate", C.c_void_p),
        ("H", C.c_int), ("W", C.c_int), ("shift", C.c_int),
        ("row_scale", C.c_void_p),
        ("precision", C.c_int), ("Bimg", C.c_void_p), ("bimg_batch_bytes", C.c_longlong),
        ("Y2", C.c_void_p), ("ldy2", C.c_int), ("n_split", C.c_int),
    ]


class ConvParams(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("ldx", C.c_int),
        ("Wt", C.c_void_p), ("ldb", C.c_int),
        ("Y", C.c_void_p), ("ldy", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("N", C.c_int),
        ("out_mode", C.c_int),
        ("R", C.c_void_p),
        ("precision", C.c_int), ("Bimg", C.c_void_p),
    ]


class MlpParams(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("ldx", C.c_int), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("W1img", C.c_void_p), ("b1", C.c_void_p), ("W2img", C.c_void_p), ("b2", C.c_void_p),
        ("res2", C.c_void_p), ("ldr2", C.c_int), ("row_scale", C.c_void_p), ("rows_per_batch", C.c_int),
        ("Y", C.c_void_p), ("ldy", C.c_int), ("M", C.c_int), ("C", C.c_int), ("hid_pad", C.c_int),
        ("precision", C.c_int),
    ]


class LocalGateParams(C.Structure):
    _fields_ = [("core_mean", C.c_void_p)] + [
        (n, C.c_void_p) for n in ("promptT", "promptb", "downT", "downb", "param", "qT", "kvT", "p2T", "p2b", "upT")
    ] + [("gate", C.c_void_p), ("B_", C.c_int), ("C", C.c_int), ("r", C.c_int)]


class WgradParams(C.Structure):
    _fields_ = [
        ("dY", C.c_void_p), ("lddy", C.c_int), ("X", C.c_void_p), ("ldx", C.c_int), ("dW", C.c_void_p),
        ("M", C.c_longlong), ("O", C.c_int), ("I", C.c_int),
        ("rows_per_batch", C.c_int), ("dw_batch_stride", C.c_longlong), ("x_row_mod", C.c_int),
        ("H", C.c_int), ("W", C.c_int), ("taps", C.c_int),
        ("so", C.c_longlong), ("si", C.c_longlong), ("st", C.c_longlong),
        ("map_mode", C.c_int), ("map_a", C.c_int), ("map_b", C.c_int),
        ("i_valid", C.c_int), ("precision", C.c_int), ("dbias", C.c_void_p),
    ]


class PackItem(C.Structure):
    _fields_ = [("W", C.c_void_p), ("img", C.c_void_p), ("ld", C.c_int), ("transposed", C.c_int), ("N", C.c_int), ("K", C.c_int)]


class LocalGateBwdWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("param", "q", "kv", "proj", "proj_bias", "up")]


EPI_BIAS, EPI_RESIDUAL, EPI_GLU, EPI_SPECTRAL, EPI_PROJ = 0, 1, 2, 3, 4
PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16 = 0, 1, 2
CONV_TOKENS, CONV_UNSHUFFLE, CONV_SHUFFLE, CONV_NCHW_RES = 0, 1, 2, 3
MAP_IDENTITY, MAP_INTERLEAVE, MAP_HALVES = 0, 1, 2

# symbol -> (restype, argtypes); also the list tests/test_abi.py checks against include/mphsir.h
_VP, _I, _LL, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
SIGNATURES = {
    "mphsir_version": (_I, []),
    "mphsir_last_error": (C.c_char_p, []),
    "mphsir_device_check": (_I, [_I, C.POINTER(_I)]),
    "mphsir_nchw_to_tokens": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "mphsir_debug_tc_counters": (None, [_VP]),
    "mphsir_debug_tc_cluster": (None, [_I]),
    "mphsir_debug_tc_psplit": (None, [_I]),
    "mphsir_gemm_plan": (_I, [_I, _I, _I, _I, _I, _I, C.POINTER(_I)]),
    "mphsir_debug_tc_reverse": (None, [_I]),
    "mphsir_debug_mlp_counters": (None, [_VP]),
    "mphsir_debug_mlp_flags": (None, [_I]),
    "mphsir_debug_dwgram_tma": (None, [_I]),
    "mphsir_debug_window_attn_tc": (None, [_I]),
    "mphsir_debug_window_attn_tc_counters": (None, [_VP]),
    "mphsir_debug_tc_tma_epilogue": (None, [_I]),
    "mphsir_debug_tc_ebox1": (None, [_I]),
    "mphsir_debug_pdl": (None, [_I]),
    "mphsir_bimg_bytes": (C.c_size_t, [_I, _I]),
    "mphsir_pack_bimg": (_I, [_VP, _I, _I, _LL, _VP, _I, _I, _I, _VP]),
    "mphsir_pack_bimg_multi": (_I, [C.POINTER(PackItem), _I, _VP]),
    "mphsir_gemm_fwd": (_I, [C.POINTER(GemmParams), _VP]),
    "mphsir_mlp_supported": (_I, [_I, _I]),
    "mphsir_mlp_fwd": (_I, [C.POINTER(MlpParams), _VP]),
    "mphsir_conv3x3_fwd": (_I, [C.POINTER(ConvParams), _VP]),
    "mphsir_window_attn_fwd": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_window_attn_band_fwd": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_window_attn_tc_supported": (_I, [_I]),
    "mphsir_window_attn_tc_enabled": (_I, []),
    "mphsir_window_attn_tc_fwd": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_dwgram_band_fwd": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_gram_reduce": (_I, [_VP, _I, _VP, _I, _I, _I, _VP]),
    "mphsir_local_gate_fwd": (_I, [C.POINTER(LocalGateParams), _VP]),
    "mphsir_local_gate_tail_fwd": (_I, [_VP, _I, C.POINTER(LocalGateParams), _VP]),
    "mphsir_local_gate2_fwd": (_I, [C.POINTER(LocalGateParams), _VP]),
    "mphsir_dwconv3x3_fwd": (_I, [_VP, _I, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_gram_partial_floats": (C.c_size_t, [_I, _I, _I, _I, C.POINTER(_I)]),
    "mphsir_gram_partial_fwd": (_I, [_VP, _I, _I, _VP, _I, _I, _VP, _I, _I, _I, _I, _VP]),
    "mphsir_gram_softmax_fwd": (_I, [_VP, _I, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "mphsir_spectral_fold_fwd": (_I, [_VP, _VP, _VP, _I, _LL, _I, _I, _I, _VP]),
    "mphsir_dwgram_supported": (_I, [_I, _I]),
    "mphsir_dwgram_partial_floats": (C.c_size_t, [_I, _I, _I, _I, _I, C.POINTER(_I)]),
    "mphsir_dwgram_fwd": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_spectral_finish_fwd": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _LL, _VP, _LL, _VP, _I, _I, _I, _VP]),
    "mphsir_tvsp_query_fwd": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "mphsir_bilinear_fwd": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_text_prompt_fwd": (_I, [_VP, _VP, _VP, _I, _I, _VP]),
    # ---- training path ----
    "mphsir_wgrad": (_I, [C.POINTER(WgradParams), _VP]),
    "mphsir_debug_wgrad_tc": (None, [_I]),
    "mphsir_wgrad_multi": (_I, [C.POINTER(WgradParams), _I, _VP]),
    "mphsir_colsum": (_I, [_VP, _I, _VP, _LL, _I, _I, _I, _I, _VP]),
    "mphsir_layernorm_fwd": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _LL, _I, _VP]),
    "mphsir_layernorm_bwd": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _I, _VP, _I, _VP, _VP, _LL, _I, _VP]),
    "mphsir_glu_bwd": (_I, [_VP, _I, _VP, _I, _LL, _I, _VP]),
    "mphsir_gdfn_gate_fwd": (_I, [_VP, _I, _VP, _I, _LL, _I, _VP]),
    "mphsir_gdfn_gate_bwd": (_I, [_VP, _I, _VP, _I, _VP, _I, _LL, _I, _VP]),
    "mphsir_axpby": (_I, [_VP, _I, _VP, _I, _LL, _I, _F, _F, _VP, _I, _I, _VP]),
    "mphsir_batch_sum": (_I, [_VP, _I, _VP, _I, _I, _LL, _I, _VP]),
    "mphsir_window_attn_bwd_groups": (_I, [_I, _I, _I, _I]),
    "mphsir_window_attn_bwd": (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_rpb_table_bwd": (_I, [_VP, _VP, _I, _VP]),
    "mphsir_window_reduce": (_I, [_VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _I, _F, _VP]),
    "mphsir_local_gate_bwd_record_ld": (_I, [_I]),
    "mphsir_local_gate_bwd": (_I, [_VP, _I, _VP, C.POINTER(LocalGateBwdWeights), _VP, _I, _I, _I, _I, _VP]),
    "mphsir_gate_apply_bwd": (_I, [_VP, _I, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_gate_apply_fwd": (_I, [_VP, _I, _VP, _I, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_spectral_bwd": (_I, [_VP, _LL, _VP, _VP, _VP, _VP, _I, _LL, _VP, _VP, _I, _I, _I, _VP]),
    "mphsir_dwconv3x3_wgrad": (_I, [_VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_pixel_unshuffle": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _I, _VP]),
    "mphsir_pixel_shuffle": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _I, _VP]),
    "mphsir_tokens_to_nchw": (_I, [_VP, _I, _VP, _I, _I, _I, _VP]),
    "mphsir_bilinear_bwd": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mphsir_tvsp_query_bwd": (_I, [_VP, _I, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "mphsir_l1_clamp_loss": (_I, [_VP, _VP, _VP, _VP, _LL, _F, _VP]),
    "mphsir_peer_window_bytes": (C.c_size_t, [_LL, _LL]),
    "mphsir_peer_window_alloc": (_I, [C.c_size_t, C.POINTER(_VP)]),
    "mphsir_peer_window_free": (_I, [_VP]),
    "mphsir_peer_export": (_I, [_VP, C.c_char_p]),
    "mphsir_peer_open": (_I, [C.c_char_p, C.POINTER(_VP)]),
    "mphsir_peer_close": (_I, [_VP]),
    "mphsir_peer_halo_exchange": (_I, [C.POINTER(_VP), _I, _I, _LL, _LL, _VP, _VP, _VP, _VP, _LL, _VP]),
    "mphsir_peer_all_reduce": (_I, [C.POINTER(_VP), _I, _I, _LL, _LL, _VP, _LL, _VP]),
    "mphsir_psnr_ssim": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "mphsir_plane_nonzero": (_I, [_VP, _I, _LL, _VP, _VP]),
    "mphsir_degrade": (_I, [_VP, _VP, _I, _I, _LL, _VP, _VP, _VP, C.c_ulonglong, _VP]),
    "mphsir_gaussian_blur": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "mphsir_sr_degrade": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "mphsir_blur2d": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "mphsir_poisson": (_I, [_VP, _VP, _VP, _I, _LL, C.c_ulonglong, _VP]),
    "mphsir_topk_mean": (_I, [_VP, _I, _LL, _I, _VP, _VP]),
    "mphsir_haze": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _LL, _VP]),
    "mphsir_degrade_structured": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, C.c_ulonglong, _VP]),
    "mphsir_adamw_step": (_I, [_VP, _VP, _VP, _VP, _LL, _F, _F, _F, _F, _F, _I, _F, _VP, _VP]),
}


def load() -> C.CDLL:
    """dlopen libmphsir.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m mp_hsir_b200.build` "
            "(there is no CPU / PyTorch fallback for the MP-HSIR hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mphsir_version() < 100:
        raise RuntimeError("libmphsir.so is older than this Python package; rebuild it")
    if os.environ.get("MPHSIR_WATC"):                  # A/B switch: tcgen05 window attention (1 | debug flags << 4) vs mma.sync (0)
        lib.mphsir_debug_window_attn_tc(int(os.environ["MPHSIR_WATC"]))
    if os.environ.get("MPHSIR_PDL") in ("0", "1"):     # A/B switch: programmatic dependent launch of the tcgen05 kernels
        lib.mphsir_debug_pdl(int(os.environ["MPHSIR_PDL"]))
    if os.environ.get("MPHSIR_PSPLIT") in ("0", "1"):  # A/B switch: few-tile GEMMs hand the passes of a row tile to several CTAs
        lib.mphsir_debug_tc_psplit(int(os.environ["MPHSIR_PSPLIT"]))
    if os.environ.get("MPHSIR_REV") in ("0", "1"):     # A/B switch: GEMM launches walk their tiles backwards
        lib.mphsir_debug_tc_reverse(int(os.environ["MPHSIR_REV"]))
    if os.environ.get("MPHSIR_PAIR") in ("0", "1"):    # A/B switch: CTA-pair (cta_group::2) instantiation of the GEMM engine
        lib.mphsir_debug_tc_cluster(int(os.environ["MPHSIR_PAIR"]))
    _lib = lib
    return lib


def last_error() -> str:
    return load().mphsir_last_error().decode(errors="replace")


def _check(rc: int, what: str) -> None:
    global LAUNCHES
    if rc != 0:
        raise RuntimeError(f"libmphsir {what} failed (rc={rc}): {last_error()}")
    LAUNCHES += 1


class Profiler:
    """Per-launch CUDA-event timing on the launching stream + algorithmic flops/bytes per kernel
    family (bench.py's roofline leg).  Off the timed path: enable only for an instrumented pass."""

    def __init__(self):
        self.records = []  # (tag, flops, bytes, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for tag, fl, by, e0, e1 in self.records:
            a = agg.setdefault(tag, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["flops"] += fl
            a["bytes"] += by
        return agg


PROFILER: Optional[Profiler] = None


def _launch(what: str, call, cost=None) -> None:
    """Run one ABI call; when PROFILER is set, bracket it with events and record its cost."""
    if PROFILER is None:
        _check(call(), what)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(call(), what)
    e1.record()
    fl, by, tag = cost() if cost is not None else (0.0, 0.0, what)
    PROFILER.records.append((tag, fl, by, e0, e1))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class View:
    """A token-major fp32 matrix living inside a (possibly wider) device buffer."""

    __slots__ = ("ptr", "ld", "rows", "cols", "keep")

    def __init__(self, ptr_: int, ld: int, rows: int, cols: int, keep=None):
        self.ptr, self.ld, self.rows, self.cols, self.keep = ptr_, ld, rows, cols, keep

    @staticmethod
    def of(t: torch.Tensor) -> "View":
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.dim() == 2
        return View(t.data_ptr(), t.shape[1], t.shape[0], t.shape[1], t)

    def cols_slice(self, c0: int, c1: int) -> "View":
        assert 0 <= c0 < c1 <= self.cols and c0 % 4 == 0
        return View(self.ptr + 4 * c0, self.ld, self.rows, c1 - c0, self.keep)

    def rows_slice(self, r0: int, r1: int) -> "View":
        return View(self.ptr + 4 * r0 * self.ld, self.ld, r1 - r0, self.cols, self.keep)

    def torch(self) -> torch.Tensor:
        """Materialise as a torch tensor (debug / tests only)."""
        base = self.keep
        off = (self.ptr - base.data_ptr()) // 4
        return base.reshape(-1).as_strided((self.rows, self.cols), (self.ld, 1), off)


# ------------------------------------------------------------------------------------------
# thin typed wrappers
# ------------------------------------------------------------------------------------------


def device_check(device: int) -> int:
    n = C.c_int(0)
    rc = load().mphsir_device_check(device, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"libmphsir device check failed: {last_error()}")
    return n.value


def nchw_to_tokens(inp: torch.Tensor, out: View) -> None:
    B, Cc, H, W = inp.shape
    _launch("nchw_to_tokens", lambda: load().mphsir_nchw_to_tokens(inp.data_ptr(), out.ptr, B, Cc, H * W, out.ld, stream_ptr()),
            lambda: (0.0, 4.0 * B * H * W * (Cc + out.ld), "nchw_to_tokens"))


class Weight:
    """A GEMM weight in both engine layouts: `bt` fp32 "in x out" [Kp, ldb] (SIMT engine) and `img`, the
    bf16 hi/lo tensor-core image of the logical [N, K] matrix (tcgen05 engine).  Either may be None.
    3-D `bt` / batched `img` ([B, bytes]) hold per-sample matrices."""

    __slots__ = ("bt", "img", "n", "k", "src")

    def __init__(self, bt: Optional[torch.Tensor], img: Optional[torch.Tensor], n: int, k: int):
        self.bt, self.img, self.n, self.k = bt, img, n, k
        self.src = None  # trainer: the [n, k_valid] parameter view the image was packed from (engine.py: L())


def bimg_bytes(n: int, k: int) -> int:
    return int(load().mphsir_bimg_bytes(n, k))


def pack_bimg(w: torch.Tensor, n: int, k: int, transposed: bool = False, img: Optional[torch.Tensor] = None):
    """Pack logical W[N,K] (2-D, or 3-D batch) into the tensor-core image (uint8 tensor).  `transposed`:
    the source is stored "in x out" (w[..., k, n])."""
    w3 = w if w.dim() == 3 else w.unsqueeze(0)
    assert w3.is_cuda and w3.dtype == torch.float32 and w3.stride(-1) == 1
    batch = w3.shape[0]
    nbytes = bimg_bytes(n, k)
    if img is None:
        img = torch.empty(batch, nbytes, device=w.device, dtype=torch.uint8)
    assert img.numel() >= batch * nbytes and img.data_ptr() % 128 == 0
    if PACK_QUEUE is not None and batch == 1:
        # deferred: flush_packs() issues one launch per 64 matrices (sources and images are kept alive by the queue)
        PACK_QUEUE.append((w3, img, w3.stride(1), int(transposed), n, k))
        return img
    _launch("pack_bimg", lambda: load().mphsir_pack_bimg(w3.data_ptr(), w3.stride(1), int(transposed),
                                                         w3.stride(0) if batch > 1 else 0, img.data_ptr(), batch, n, k,
                                                         stream_ptr()),
            lambda: (0.0, 4.0 * batch * n * k + batch * nbytes, "pack_bimg"))
    return img


PACK_QUEUE: Optional[list] = None  # set to [] to defer pack_bimg calls; flush_packs() launches them in bulk


def flush_packs() -> None:
    """launch every deferred weight-image pack (mphsir_pack_bimg_multi) and leave deferred mode"""
    global PACK_QUEUE
    q, PACK_QUEUE = PACK_QUEUE, None
    if not q:
        return
    arr = (PackItem * len(q))()
    nbytes = 0.0
    for j, (w3, img, ld, tr, n, k) in enumerate(q):
        arr[j] = PackItem(w3.data_ptr(), img.data_ptr(), ld, tr, n, k)
        nbytes += 4.0 * n * k + img.numel()
    _launch("pack_bimg_multi", lambda: load().mphsir_pack_bimg_multi(arr, len(q), stream_ptr()), lambda: (0.0, nbytes, "pack_bimg"))
    del q


def gemm(A: View, Bt, Y: View, N: int, *, K: Optional[int] = None, ln=None, bias=None, precision: int = 0,
         epi: int = EPI_BIAS, res1: Optional[View] = None, res2: Optional[View] = None,
         gsrc: Optional[View] = None, gate: Optional[torch.Tensor] = None, H: int = 0, W: int = 0,
         shift: int = 0, rows_per_batch: int = 0, b_batch_stride: int = 0, a_row_mod: int = 0,
         M: Optional[int] = None, row_scale: Optional[torch.Tensor] = None, Y2: Optional[View] = None,
         n_split: int = 0) -> None:
    p = GemmParams()
    p.A, p.lda, p.a_row_mod = A.ptr, A.ld, a_row_mod
    if Y2 is not None:
        p.Y2, p.ldy2, p.n_split = Y2.ptr, Y2.ld, n_split
    if isinstance(Bt, Weight):
        wobj = Bt
        if precision == PREC_FP32_SIMT:
            Bt = wobj.bt
        else:
            Bt = None
            p.Bimg = wobj.img.data_ptr()
            if wobj.img.dim() == 2 and wobj.img.shape[0] > 1:
                p.bimg_batch_bytes = wobj.img.stride(0)
                b_batch_stride = 1  # marks per-sample weights for the cost model below
    if Bt is not None:
        p.Bt, p.ldb = Bt.data_ptr(), Bt.shape[-1]
        if Bt.dim() == 3 and b_batch_stride == 0:
            b_batch_stride = Bt.stride(0)
    p.precision = precision
    p.b_batch_stride, p.rows_per_batch = (b_batch_stride if Bt is not None else 0), rows_per_batch
    p.Y, p.ldy = Y.ptr, Y.ld
    p.M, p.N, p.K = (A.rows if M is None else M), N, (A.cols if K is None else K)
    if ln is not None:
        p.ln_gamma, p.ln_beta = ln[0].data_ptr(), ln[1].data_ptr()
    p.bias = ptr(bias)
    p.epi = epi
    if res1 is not None:
        p.res1, p.ldr1 = res1.ptr, res1.ld
    if res2 is not None:
        p.res2, p.ldr2 = res2.ptr, res2.ld
    if gsrc is not None:
        p.gsrc, p.ldg = gsrc.ptr, gsrc.ld
    p.gate = ptr(gate)
    p.H, p.W, p.shift = H, W, shift
    p.row_scale = ptr(row_scale)
    def cost():
        m, n, k = p.M, p.N, p.K
        n_out = n // 2 if epi == EPI_GLU else n
        reads = m * k + k * n + sum(m * n for t in (res1, res2, gsrc) if t is not None)
        tag = ("gemm", "gemm_tc3", "gemm_tc1")[precision] + ("+ln" if ln is not None else "") + ("", "+res", "+glu", "+spectral", "+proj")[epi]
        return 2.0 * m * n * k, 4.0 * (reads + m * n_out), tag

    _launch("gemm_fwd", lambda: load().mphsir_gemm_fwd(C.byref(p), stream_ptr()), cost)


def mlp_supported(Cc: int, hid_pad: int) -> bool:
    return bool(load().mphsir_mlp_supported(Cc, hid_pad))


def mlp(X: View, ln, W1: "Weight", b1: torch.Tensor, W2: "Weight", b2: torch.Tensor, Y: View, hid_pad: int,
        precision: int, res2: Optional[View] = None, row_scale: Optional[torch.Tensor] = None,
        rows_per_batch: int = 0) -> None:
    """Fused LN -> fc1 -> value*gelu(gate) -> fc2 -> + residual(s) (one tcgen05 kernel)."""
    p = MlpParams()
    p.X, p.ldx = X.ptr, X.ld
    p.ln_gamma, p.ln_beta = ln[0].data_ptr(), ln[1].data_ptr()
    p.W1img, p.b1, p.W2img, p.b2 = W1.img.data_ptr(), b1.data_ptr(), W2.img.data_ptr(), b2.data_ptr()
    if res2 is not None:
        p.res2, p.ldr2 = res2.ptr, res2.ld
    p.row_scale, p.rows_per_batch = ptr(row_scale), rows_per_batch
    p.Y, p.ldy = Y.ptr, Y.ld
    p.M, p.C, p.hid_pad, p.precision = X.rows, X.cols, hid_pad, precision
    m, c = X.rows, X.cols
    _launch("mlp_fwd", lambda: load().mphsir_mlp_fwd(C.byref(p), stream_ptr()),
            lambda: (2.0 * m * c * 3 * hid_pad, 4.0 * m * c * (3 if res2 is None else 4), ("", "mlp_tc3", "mlp_tc1")[precision]))


def conv3x3(X: View, Wt, Y_ptr: int, ldy: int, B: int, H: int, W: int, Cin: int, N: int,
            out_mode: int = CONV_TOKENS, R: Optional[torch.Tensor] = None, precision: int = 0) -> None:
    p = ConvParams()
    p.X, p.ldx = X.ptr, X.ld
    if isinstance(Wt, Weight):
        if precision == PREC_FP32_SIMT:
            p.Wt, p.ldb = Wt.bt.data_ptr(), Wt.bt.shape[-1]
        else:
            p.Bimg = Wt.img.data_ptr()
    else:
        p.Wt, p.ldb = Wt.data_ptr(), Wt.shape[-1]
    p.precision = precision
    p.Y, p.ldy = Y_ptr, ldy
    p.B, p.H, p.W, p.Cin, p.N = B, H, W, Cin, N
    p.out_mode = out_mode
    p.R = ptr(R)
    m = B * H * W
    _launch("conv3x3_fwd", lambda: load().mphsir_conv3x3_fwd(C.byref(p), stream_ptr()),
            lambda: (2.0 * m * N * 9 * Cin, 4.0 * (m * Cin + m * N * (2 if R is not None else 1) + 9 * Cin * N),
                     ("conv3x3", "conv3x3_tc3", "conv3x3_tc1")[precision]))


def window_attn(qkv: View, bias: torch.Tensor, out: View, win_mean: torch.Tensor, B: int, H: int, W: int,
                Cc: int, heads: int, shift: int, precision: int = 0, mask_H: Optional[int] = None, mask_y0: int = 0,
                bias_t: Optional[torch.Tensor] = None) -> None:
    """mask_H / mask_y0: row band of a sharded scene (mphsir_window_attn_band_fwd); default = the whole image.
    bias_t: the bias transposed to [heads, key, query] — when given (and the head dim is 32 / 64, tensor-core precision) the
    TMA-fed tcgen05 kernel runs instead of the mma.sync one."""
    n = B * H * W
    mH = H if mask_H is None else mask_H
    L = load()
    if bias_t is not None and precision != PREC_FP32_SIMT and L.mphsir_window_attn_tc_supported(Cc // heads) and L.mphsir_window_attn_tc_enabled():
        _launch("window_attn_tc_fwd",
                lambda: L.mphsir_window_attn_tc_fwd(qkv.ptr, qkv.ld, bias_t.data_ptr(), out.ptr, out.ld, win_mean.data_ptr(), B, H, W,
                                                    Cc, heads, shift, precision, mH, mask_y0, stream_ptr()),
                lambda: (4.0 * n * 64 * Cc, 4.0 * (4 * n * Cc + n // 64 * Cc), ("", "window_attn_tc3", "window_attn_tc1")[precision]))
        return
    _launch("window_attn_fwd",
            lambda: L.mphsir_window_attn_band_fwd(qkv.ptr, qkv.ld, bias.data_ptr(), out.ptr, out.ld,
                                                  win_mean.data_ptr(), B, H, W, Cc, heads, shift, precision, mH, mask_y0,
                                                  stream_ptr()),
            lambda: (4.0 * n * 64 * Cc, 4.0 * (4 * n * Cc + n // 64 * Cc), ("window_attn", "window_attn_mma3", "window_attn_mma1")[precision]))


def local_gate(core_mean: torch.Tensor, w: dict, gate: torch.Tensor, B_: int, Cc: int, r: int) -> None:
    p = LocalGateParams()
    p.core_mean = core_mean.data_ptr()
    for n in ("promptT", "promptb", "downT", "downb", "param", "qT", "kvT", "p2T", "p2b", "upT"):
        setattr(p, n, w[n].data_ptr())
    p.gate, p.B_, p.C, p.r = gate.data_ptr(), B_, Cc, r
    _launch("local_gate_fwd", lambda: load().mphsir_local_gate_fwd(C.byref(p), stream_ptr()),
            lambda: (2.0 * B_ * (Cc * 128 + 2 * Cc * r + 128 * r), 8.0 * B_ * Cc, "local_gate"))


def local_gate2(core_mean: torch.Tensor, w: dict, gate: torch.Tensor, B_: int, Cc: int, r: int) -> None:
    """One-launch local spectral gate (net/MP_HSIR.py:132-152 on the window means; proj folded into promptT / downT)."""
    p = LocalGateParams()
    p.core_mean = core_mean.data_ptr()
    for n in ("promptT", "promptb", "downT", "downb", "param", "qT", "kvT", "p2T", "p2b", "upT"):
        setattr(p, n, w[n].data_ptr())
    p.gate, p.B_, p.C, p.r = gate.data_ptr(), B_, Cc, r
    _launch("local_gate2_fwd", lambda: load().mphsir_local_gate2_fwd(C.byref(p), stream_ptr()),
            lambda: (2.0 * B_ * (Cc * 128 + 2 * Cc * r + 128 * r), 8.0 * B_ * Cc, "local_gate2"))


def local_gate_tail(logits: View, w: dict, gate: torch.Tensor, B_: int, Cc: int, r: int) -> None:
    """r-sized remainder of the local spectral gate; `logits` [B_, 128 + r] comes from one GEMM."""
    p = LocalGateParams()
    for n in ("param", "qT", "kvT", "p2T", "p2b", "upT"):
        setattr(p, n, w[n].data_ptr())
    p.gate, p.B_, p.C, p.r = gate.data_ptr(), B_, Cc, r
    _launch("local_gate_tail_fwd", lambda: load().mphsir_local_gate_tail_fwd(logits.ptr, logits.ld, C.byref(p), stream_ptr()),
            lambda: (2.0 * B_ * (128 * r + 5 * r * r + Cc * r), 4.0 * B_ * (Cc + 128 + r), "local_gate_tail"))


def dwconv3x3(X: View, w9: torch.Tensor, Y: View, B: int, H: int, W: int, Cc: int, gate_half: int = 0) -> None:
    n = B * H * W
    _launch("dwconv3x3_fwd",
            lambda: load().mphsir_dwconv3x3_fwd(X.ptr, X.ld, w9.data_ptr(), Y.ptr, Y.ld, B, H, W, Cc, gate_half,
                                                stream_ptr()),
            lambda: (18.0 * n * Cc, 4.0 * n * (Cc + (gate_half if gate_half else Cc)),
                     "dwconv3x3+gate" if gate_half else "dwconv3x3"))


def gram_partial_floats(B: int, heads: int, c: int, HW: int):
    n = C.c_int(0)
    f = load().mphsir_gram_partial_floats(B, heads, c, HW, C.byref(n))
    return int(f), n.value


def gram_partial(q: View, q_shared: bool, k: View, k_shared: bool, partial: torch.Tensor, B: int, HW: int,
                 heads: int, c: int) -> None:
    _launch("gram_partial_fwd",
            lambda: load().mphsir_gram_partial_fwd(q.ptr, q.ld, int(q_shared), k.ptr, k.ld, int(k_shared),
                                                   partial.data_ptr(), B, HW, heads, c, stream_ptr()),
            lambda: (2.0 * B * HW * heads * c * (c + 2), 8.0 * B * HW * heads * c, "gram_partial"))


def gram_softmax(partial: torch.Tensor, n_chunks: int, temperature: torch.Tensor, attn: torch.Tensor, B: int,
                 heads: int, c: int, scratch: Optional[torch.Tensor] = None) -> None:
    if scratch is None:
        scratch = torch.empty(B * heads * (c * c + 2 * c), device=partial.device, dtype=torch.float32)
    _launch("gram_softmax_fwd",
            lambda: load().mphsir_gram_softmax_fwd(partial.data_ptr(), n_chunks, temperature.data_ptr(),
                                                   attn.data_ptr(), scratch.data_ptr(), B, heads, c, stream_ptr()),
            lambda: (0.0, 4.0 * B * heads * (n_chunks + 1) * c * c, "gram_softmax"))


def spectral_fold(attn: torch.Tensor, WoutT: torch.Tensor, Mt: torch.Tensor, B: int, heads: int, c: int) -> None:
    # Mt: [B, Cp, ldm]
    _launch("spectral_fold_fwd",
            lambda: load().mphsir_spectral_fold_fwd(attn.data_ptr(), WoutT.data_ptr(), Mt.data_ptr(), Mt.shape[2],
                                                    Mt.shape[1] * Mt.shape[2], B, heads, c, stream_ptr()),
            lambda: (2.0 * B * heads * c * c * heads * c, 4.0 * B * (heads * c) ** 2, "spectral_fold"))


def dwgram_supported(Cc: int, c: int) -> bool:
    return bool(load().mphsir_dwgram_supported(Cc, c))


def dwgram_partial_floats(B: int, heads: int, c: int, H: int, W: int):
    n = C.c_int(0)
    f = load().mphsir_dwgram_partial_floats(B, heads, c, H, W, C.byref(n))
    return int(f), n.value


def dwgram(X: View, w9: torch.Tensor, Vout: View, partial: torch.Tensor, B: int, H: int, W: int, Cc: int, heads: int,
           precision: int, gram_rows: Optional[Tuple[int, int]] = None) -> None:
    """gram_rows: (y0, y1) tile-aligned row range whose tiles enter the Gram statistics (row band of a sharded scene);
    default = every row."""
    n = B * H * W
    gy0, gy1 = (0, H) if gram_rows is None else gram_rows
    _launch("dwgram_fwd",
            lambda: load().mphsir_dwgram_band_fwd(X.ptr, X.ld, w9.data_ptr(), Vout.ptr, Vout.ld, partial.data_ptr(), B, H, W,
                                                  Cc, heads, precision, gy0, gy1, stream_ptr()),
            lambda: (2.0 * n * (27 * Cc + Cc * (Cc // heads + 2)), 16.0 * n * Cc, ("", "dwgram3", "dwgram1")[precision]))


def gram_reduce(partial: torch.Tensor, n_chunks: int, reduced: torch.Tensor, B: int, heads: int, c: int) -> None:
    """reduced[B*heads, c*c+2c] = sum over the n_chunks partials (the per-rank statistics of a sharded scene)"""
    _launch("gram_reduce", lambda: load().mphsir_gram_reduce(partial.data_ptr(), n_chunks, reduced.data_ptr(), B, heads, c,
                                                             stream_ptr()),
            lambda: (0.0, 4.0 * B * heads * (n_chunks + 1) * (c * c + 2 * c), "gram_reduce"))


def spectral_finish(partial: torch.Tensor, n_chunks: int, scratch: torch.Tensor, temperature: torch.Tensor,
                    WoutT: torch.Tensor, Mt: Optional[torch.Tensor], img: Optional[torch.Tensor], B: int, heads: int,
                    c: int, attn_out: Optional[torch.Tensor] = None) -> None:
    """gram reduce + softmax + project_out fold (+ tensor-core image) in one ABI call."""
    ldm = Mt.shape[2] if Mt is not None else 0
    mstride = Mt.shape[1] * Mt.shape[2] if Mt is not None else 0
    istride = img.stride(0) if img is not None else 0
    _launch("spectral_finish_fwd",
            lambda: load().mphsir_spectral_finish_fwd(partial.data_ptr(), n_chunks, scratch.data_ptr(),
                                                      temperature.data_ptr(), WoutT.data_ptr(), ptr(Mt), ldm, mstride,
                                                      ptr(img), istride, ptr(attn_out), B, heads, c, stream_ptr()),
            lambda: (2.0 * B * heads * c * c * heads * c, 4.0 * B * heads * (n_chunks + 1) * c * c, "spectral_finish"))


def tvsp_query(clip_b: torch.Tensor, weights: torch.Tensor, learnable: torch.Tensor, Q: View, B: int, T: int,
               D: int, ps: int) -> None:
    _launch("tvsp_query_fwd",
            lambda: load().mphsir_tvsp_query_fwd(clip_b.data_ptr(), weights.data_ptr(), learnable.data_ptr(), Q.ptr,
                                                 B, T, D, ps, stream_ptr()),
            lambda: (0.0, 4.0 * B * ps * ps * D, "tvsp_query"))


def bilinear(X: View, Y: View, B: int, h: int, w: int, H: int, W: int, Cc: int) -> None:
    _launch("bilinear_fwd",
            lambda: load().mphsir_bilinear_fwd(X.ptr, X.ld, Y.ptr, Y.ld, B, h, w, H, W, Cc, stream_ptr()),
            lambda: (0.0, 4.0 * B * Cc * (h * w + H * W), "bilinear"))


def text_prompt(weights: torch.Tensor, clip: torch.Tensor, clip_b: torch.Tensor, B: int, T: int) -> None:
    _launch("text_prompt_fwd",
            lambda: load().mphsir_text_prompt_fwd(weights.data_ptr(), clip.data_ptr(), clip_b.data_ptr(), B, T,
                                                  stream_ptr()),
            lambda: (0.0, 4.0 * B * 512, "text_prompt"))


# ------------------------------------------------------------------------------------------
# training path (include/mphsir.h, "Training path")
# ------------------------------------------------------------------------------------------


def _wgrad_params(dY: View, X: View, dW: torch.Tensor, precision: int, *, M: Optional[int] = None, so: Optional[int] = None,
                  si: int = 1, st: int = 0, map_mode: int = MAP_IDENTITY, map_a: Optional[int] = None, map_b: int = 0,
                  i_valid: int = 0, rows_per_batch: int = 0, dw_batch_stride: int = 0, x_row_mod: int = 0, taps: int = 0,
                  H: int = 0, W: int = 0, dw_offset: int = 0, dbias: Optional[torch.Tensor] = None) -> WgradParams:
    p = WgradParams()
    p.dY, p.lddy, p.X, p.ldx = dY.ptr, dY.ld, X.ptr, X.ld
    p.dW = dW.data_ptr() + 4 * dw_offset
    p.M, p.O, p.I = (dY.rows if M is None else M), dY.cols, X.cols
    p.rows_per_batch, p.dw_batch_stride, p.x_row_mod = rows_per_batch, dw_batch_stride, x_row_mod
    p.H, p.W, p.taps = H, W, taps
    iv = i_valid if i_valid > 0 else X.cols
    p.so, p.si, p.st = (iv * si if so is None else so), si, st
    p.map_mode, p.map_a, p.map_b = map_mode, (dY.cols if map_a is None else map_a), map_b
    p.i_valid, p.precision = iv, precision
    p.dbias = ptr(dbias)
    return p


def wgrad(dY: View, X: View, dW: torch.Tensor, precision: int, **kw) -> None:
    """dW[map(o)*so + i*si + tap*st] += sum_m dY[m,o] * X[src(m), i]  (dW: flat fp32 gradient storage);
    dbias[map(o)] += sum_m dY[m,o] when given (plain mode).  Keywords: see _wgrad_params."""
    p = _wgrad_params(dY, X, dW, precision, **kw)
    m, o, i, taps = p.M, p.O, p.I, p.taps
    _launch("wgrad", lambda: load().mphsir_wgrad(C.byref(p), stream_ptr()),
            lambda: (2.0 * m * o * i * max(taps, 1), 4.0 * m * (o + i) * max(taps, 1), ("", "wgrad3", "wgrad1")[precision]))


def wgrad_multi(problems, precision: int) -> None:
    """several small plain-mode weight gradients in one launch; problems: [(dY, X, dW, kwargs), ...] (at most 8)."""
    arr = (WgradParams * len(problems))()
    fl = by = 0.0
    for j, (dY, X, dW, kw) in enumerate(problems):
        arr[j] = _wgrad_params(dY, X, dW, precision, **kw)
        fl += 2.0 * arr[j].M * arr[j].O * arr[j].I
        by += 4.0 * arr[j].M * (arr[j].O + arr[j].I)
    _launch("wgrad_multi", lambda: load().mphsir_wgrad_multi(arr, len(problems), stream_ptr()),
            lambda: (fl, by, ("", "wgrad_multi3", "wgrad_multi1")[precision]))


def colsum(X: View, out: torch.Tensor, map_mode: int = MAP_IDENTITY, map_a: Optional[int] = None, map_b: int = 0,
           M: Optional[int] = None) -> None:
    m = X.rows if M is None else M
    _launch("colsum", lambda: load().mphsir_colsum(X.ptr, X.ld, out.data_ptr(), m, X.cols, map_mode,
                                                   X.cols if map_a is None else map_a, map_b, stream_ptr()),
            lambda: (0.0, 4.0 * m * X.cols, "colsum"))


def layernorm_fwd(X: View, ln, Y: View, stats: Optional[torch.Tensor]) -> None:
    _launch("layernorm_fwd", lambda: load().mphsir_layernorm_fwd(X.ptr, X.ld, ln[0].data_ptr(), ln[1].data_ptr(), Y.ptr, Y.ld,
                                                                 ptr(stats), X.rows, X.cols, stream_ptr()),
            lambda: (0.0, 8.0 * X.rows * X.cols, "layernorm_fwd"))


def layernorm_bwd(X: View, stats: torch.Tensor, gamma: torch.Tensor, G: View, add: Optional[View], dX: View,
                  dgamma: torch.Tensor, dbeta: torch.Tensor) -> None:
    _launch("layernorm_bwd",
            lambda: load().mphsir_layernorm_bwd(X.ptr, X.ld, stats.data_ptr(), gamma.data_ptr(), G.ptr, G.ld,
                                                add.ptr if add is not None else None, add.ld if add is not None else 0,
                                                dX.ptr, dX.ld, dgamma.data_ptr(), dbeta.data_ptr(), X.rows, X.cols, stream_ptr()),
            lambda: (0.0, 4.0 * X.rows * X.cols * (3 if add is None else 4), "layernorm_bwd"))


def glu_bwd(H: View, dHid: View, hid_pad: int) -> None:
    _launch("glu_bwd", lambda: load().mphsir_glu_bwd(H.ptr, H.ld, dHid.ptr, dHid.ld, H.rows, hid_pad, stream_ptr()),
            lambda: (0.0, 4.0 * H.rows * 6 * hid_pad, "glu_bwd"))


def gdfn_gate_fwd(T: View, Y: View, hid_pad: int) -> None:
    _launch("gdfn_gate_fwd", lambda: load().mphsir_gdfn_gate_fwd(T.ptr, T.ld, Y.ptr, Y.ld, T.rows, hid_pad, stream_ptr()),
            lambda: (0.0, 4.0 * T.rows * 3 * hid_pad, "gdfn_gate_fwd"))


def gdfn_gate_bwd(T: View, dY: View, dT: View, hid_pad: int) -> None:
    _launch("gdfn_gate_bwd", lambda: load().mphsir_gdfn_gate_bwd(T.ptr, T.ld, dY.ptr, dY.ld, dT.ptr, dT.ld, T.rows, hid_pad,
                                                                 stream_ptr()),
            lambda: (0.0, 4.0 * T.rows * 5 * hid_pad, "gdfn_gate_bwd"))


def axpby(X: View, Y: View, alpha: float = 1.0, beta: float = 0.0, row_scale: Optional[torch.Tensor] = None,
          rows_per_batch: int = 0, x_row_mod: int = 0, M: Optional[int] = None) -> None:
    m = Y.rows if M is None else M
    _launch("axpby", lambda: load().mphsir_axpby(X.ptr, X.ld, Y.ptr, Y.ld, m, Y.cols, alpha, beta, ptr(row_scale),
                                                 rows_per_batch, x_row_mod, stream_ptr()),
            lambda: (0.0, 4.0 * m * Y.cols * (2 if beta == 0.0 else 3), "axpby"))


def batch_sum(X: View, Y: View, B: int) -> None:
    _launch("batch_sum", lambda: load().mphsir_batch_sum(X.ptr, X.ld, Y.ptr, Y.ld, B, Y.rows, Y.cols, stream_ptr()),
            lambda: (0.0, 4.0 * (B + 1) * Y.rows * Y.cols, "batch_sum"))


def window_attn_bwd_groups(B: int, H: int, W: int, heads: int) -> int:
    return int(load().mphsir_window_attn_bwd_groups(B, H, W, heads))


def window_attn_bwd(qkv: View, bias: torch.Tensor, dO: View, dqkv: View, partial: torch.Tensor, groups: int, B: int,
                    H: int, W: int, Cc: int, heads: int, shift: int, precision: int = 0) -> None:
    n = B * H * W
    _launch("window_attn_bwd",
            lambda: load().mphsir_window_attn_bwd(qkv.ptr, qkv.ld, bias.data_ptr(), dO.ptr, dO.ld, dqkv.ptr, dqkv.ld,
                                                  partial.data_ptr(), groups, B, H, W, Cc, heads, shift, precision, stream_ptr()),
            lambda: (10.0 * n * 64 * Cc, 4.0 * 7 * n * Cc, ("window_attn_bwd", "window_attn_bwd_mma3", "window_attn_bwd_mma1")[precision]))


def rpb_table_bwd(dbias: torch.Tensor, dtable: torch.Tensor, heads: int) -> None:
    _launch("rpb_table_bwd", lambda: load().mphsir_rpb_table_bwd(dbias.data_ptr(), dtable.data_ptr(), heads, stream_ptr()))


def window_reduce(A: View, Bm: Optional[View], out: torch.Tensor, B: int, H: int, W: int, Cc: int, shift: int,
                  scale: float) -> None:
    _launch("window_reduce",
            lambda: load().mphsir_window_reduce(A.ptr, A.ld, Bm.ptr if Bm is not None else None, Bm.ld if Bm is not None else 0,
                                                out.data_ptr(), B, H, W, Cc, shift, scale, stream_ptr()),
            lambda: (0.0, 4.0 * B * H * W * Cc * (1 if Bm is None else 2), "window_reduce"))


def local_gate_bwd_record_ld(r: int) -> int:
    return int(load().mphsir_local_gate_bwd_record_ld(r))


def local_gate_bwd(LL: View, dG: torch.Tensor, w: dict, record: View, B_: int, Cc: int, r: int) -> None:
    ws = LocalGateBwdWeights()
    for n in ("param", "q", "kv", "proj", "proj_bias", "up"):
        setattr(ws, n, w[n].data_ptr())
    _launch("local_gate_bwd", lambda: load().mphsir_local_gate_bwd(LL.ptr, LL.ld, dG.data_ptr(), C.byref(ws), record.ptr,
                                                                   record.ld, B_, Cc, r, stream_ptr()))


def gate_apply_fwd(X: View, SA: View, gate: torch.Tensor, row_scale: Optional[torch.Tensor], U: View, B: int, H: int, W: int,
                   Cc: int, shift: int) -> None:
    _launch("gate_apply_fwd", lambda: load().mphsir_gate_apply_fwd(X.ptr, X.ld, SA.ptr, SA.ld, gate.data_ptr(), ptr(row_scale), U.ptr,
                                                                   U.ld, B, H, W, Cc, shift, stream_ptr()),
            lambda: (0.0, 12.0 * B * H * W * Cc, "gate_apply_fwd"))


def gate_apply_bwd(dU: View, gate: torch.Tensor, dMean: torch.Tensor, dSA: View, B: int, H: int, W: int, Cc: int,
                   shift: int) -> None:
    _launch("gate_apply_bwd", lambda: load().mphsir_gate_apply_bwd(dU.ptr, dU.ld, gate.data_ptr(), dMean.data_ptr(), dSA.ptr,
                                                                   dSA.ld, B, H, W, Cc, shift, stream_ptr()),
            lambda: (0.0, 8.0 * B * H * W * Cc, "gate_apply_bwd"))


def spectral_bwd(P: torch.Tensor, Wout: torch.Tensor, gsum: torch.Tensor, temperature: torch.Tensor, Wb: torch.Tensor,
                 dWout: torch.Tensor, dTemp: torch.Tensor, B: int, heads: int, c: int) -> None:
    """P [B,C,C]; Wb [B,2C,ldwb]."""
    Cc = heads * c
    _launch("spectral_bwd",
            lambda: load().mphsir_spectral_bwd(P.data_ptr(), Cc * Cc, Wout.data_ptr(), gsum.data_ptr(), temperature.data_ptr(),
                                               Wb.data_ptr(), Wb.shape[2], Wb.shape[1] * Wb.shape[2], dWout.data_ptr(),
                                               dTemp.data_ptr(), B, heads, c, stream_ptr()),
            lambda: (4.0 * B * heads * c * c * Cc, 4.0 * B * (Cc * Cc + 4 * Cc * Cc), "spectral_bwd"))


def dwconv3x3_wgrad(X: View, dY: View, dW: torch.Tensor, B: int, H: int, W: int, Cc: int, map_mode: int = MAP_IDENTITY,
                    map_a: Optional[int] = None, map_b: int = 0) -> None:
    """dW: gradient of the reference depthwise weight [C,1,3,3] (+=)."""
    _launch("dwconv3x3_wgrad", lambda: load().mphsir_dwconv3x3_wgrad(X.ptr, X.ld, dY.ptr, dY.ld, dW.data_ptr(), B, H, W, Cc,
                                                                     map_mode, Cc if map_a is None else map_a, map_b,
                                                                     stream_ptr()),
            lambda: (18.0 * B * H * W * Cc, 8.0 * B * H * W * Cc, "dwconv3x3_wgrad"))


def pixel_unshuffle(X: View, Y: View, B: int, H: int, W: int, Cc: int) -> None:
    _launch("pixel_unshuffle", lambda: load().mphsir_pixel_unshuffle(X.ptr, X.ld, Y.ptr, Y.ld, B, H, W, Cc, stream_ptr()),
            lambda: (0.0, 8.0 * B * H * W * Cc, "pixel_shuffle"))


def pixel_shuffle(X: View, Y: View, B: int, H: int, W: int, Cc: int) -> None:
    _launch("pixel_shuffle", lambda: load().mphsir_pixel_shuffle(X.ptr, X.ld, Y.ptr, Y.ld, B, H, W, Cc, stream_ptr()),
            lambda: (0.0, 32.0 * B * H * W * Cc, "pixel_shuffle"))


def tokens_to_nchw(X: View, out: torch.Tensor, B: int, Cc: int, HW: int) -> None:
    _launch("tokens_to_nchw", lambda: load().mphsir_tokens_to_nchw(X.ptr, X.ld, out.data_ptr(), B, Cc, HW, stream_ptr()),
            lambda: (0.0, 8.0 * B * Cc * HW, "tokens_to_nchw"))


def bilinear_bwd(dY: View, dX: View, B: int, h: int, w: int, H: int, W: int, Cc: int) -> None:
    _launch("bilinear_bwd", lambda: load().mphsir_bilinear_bwd(dY.ptr, dY.ld, dX.ptr, dX.ld, B, h, w, H, W, Cc, stream_ptr()))


def tvsp_query_bwd(dQ: View, clip_b: torch.Tensor, weights: torch.Tensor, dLearn: torch.Tensor, B: int, T: int, D: int,
                   ps: int) -> None:
    _launch("tvsp_query_bwd", lambda: load().mphsir_tvsp_query_bwd(dQ.ptr, dQ.ld, clip_b.data_ptr(), weights.data_ptr(),
                                                                   dLearn.data_ptr(), B, T, D, ps, stream_ptr()))


def l1_clamp_loss(out: torch.Tensor, clean: torch.Tensor, dOut: torch.Tensor, loss: torch.Tensor, grad_scale: float = 1.0) -> None:
    _launch("l1_clamp_loss", lambda: load().mphsir_l1_clamp_loss(out.data_ptr(), clean.data_ptr(), dOut.data_ptr(),
                                                                 loss.data_ptr(), out.numel(), grad_scale, stream_ptr()),
            lambda: (0.0, 12.0 * out.numel(), "l1_clamp_loss"))


def adamw_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, beta1: float, beta2: float,
               eps: float, weight_decay: float, step: int, grad_scale: float = 1.0, dyn: Optional[torch.Tensor] = None) -> None:
    _launch("adamw_step", lambda: load().mphsir_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                                           lr, beta1, beta2, eps, weight_decay, step, grad_scale, ptr(dyn),
                                                           stream_ptr()),
            lambda: (0.0, 28.0 * p.numel(), "adamw"))


def psnr_ssim_sums(restored: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 pair -> float64 [B*C, 2] {sum of squared error, sum of the 7x7-window SSIM index} per band plane"""
    assert restored.shape == clean.shape and restored.dim() == 4 and restored.is_cuda and clean.is_cuda
    r, c = restored.detach().float().contiguous(), clean.detach().float().contiguous()
    B, Cc, H, W = r.shape
    sums = torch.empty(B * Cc, 2, dtype=torch.float64, device=r.device)
    _launch("psnr_ssim", lambda: load().mphsir_psnr_ssim(r.data_ptr(), c.data_ptr(), B * Cc, H, W, sums.data_ptr(), stream_ptr()),
            lambda: (0.0, 8.0 * r.numel(), "psnr_ssim"))
    return sums


def plane_nonzero(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> int32 [B*C]: 1 where the band plane has any non-zero element"""
    xx = x.detach().float().contiguous()
    B, Cc, H, W = xx.shape
    out = torch.empty(B * Cc, dtype=torch.int32, device=xx.device)
    _launch("plane_nonzero", lambda: load().mphsir_plane_nonzero(xx.data_ptr(), B * Cc, H * W, out.data_ptr(), stream_ptr()),
            lambda: (0.0, 4.0 * xx.numel(), "plane_nonzero"))
    return out


def degrade_structured(x: torch.Tensor, colmul: torch.Tensor, coladd: torch.Tensor, impulse: torch.Tensor, active: torch.Tensor,
                       seed: int) -> None:
    """in place: x = x * colmul[b,c,col] + coladd[b,c,col], then impulse flips (utils/degradation_utils.py:41-84) for active samples"""
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32 and colmul.numel() == B * Cc * W and coladd.numel() == B * Cc * W
    assert impulse.numel() == B * Cc and active.numel() == B and active.dtype == torch.int32
    _launch("degrade_structured", lambda: load().mphsir_degrade_structured(x.data_ptr(), B, Cc, H, W, colmul.data_ptr(), coladd.data_ptr(),
                                                                           impulse.data_ptr(), active.data_ptr(),
                                                                           seed & 0xFFFFFFFFFFFFFFFF, stream_ptr()),
            lambda: (0.0, 8.0 * x.numel(), "degrade_structured"))


def gaussian_blur(x: torch.Tensor, out: torch.Tensor, ksize: torch.Tensor, kmax: int) -> None:
    """out[b] = every band of x[b] blurred with the k x k Gaussian of utils/degradation_utils.py:91-108 where ksize[b] > 0;
    other samples' planes of `out` are left as they are.  ksize: int32 [B] on the device."""
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape and x.dtype == torch.float32
    assert ksize.dtype == torch.int32 and ksize.numel() == B and ksize.is_cuda
    _launch("gaussian_blur", lambda: load().mphsir_gaussian_blur(x.data_ptr(), out.data_ptr(), ksize.data_ptr(), B, Cc, H, W, kmax,
                                                                 stream_ptr()),
            lambda: (0.0, 8.0 * x.numel(), "gaussian_blur"))


def sr_degrade(x: torch.Tensor, out: torch.Tensor, factor: torch.Tensor) -> None:
    """out[b] = every band of x[b] bicubically down-sampled by factor[b] and replicated back (utils/degradation_utils.py:165-176,
    :189-200) where factor[b] > 0; other samples' planes of `out` are left as they are.  factor: int32 [B] on the device."""
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape and x.dtype == torch.float32
    assert factor.dtype == torch.int32 and factor.numel() == B and factor.is_cuda
    _launch("sr_degrade", lambda: load().mphsir_sr_degrade(x.data_ptr(), out.data_ptr(), factor.data_ptr(), B, Cc, H, W, stream_ptr()),
            lambda: (0.0, 8.0 * x.numel(), "sr_degrade"))


def blur2d(x: torch.Tensor, out: torch.Tensor, taps: torch.Tensor, active: torch.Tensor) -> None:
    """out[b] = every band of x[b] cross-correlated with the k x k kernel `taps` (zero padding k // 2) where active[b] != 0 — the
    circle / square / motion blur of utils/degradation_utils.py:110-163; other samples' planes of `out` are left as they are."""
    B, Cc, H, W = x.shape
    k = taps.shape[-1]
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape and x.dtype == torch.float32
    assert taps.is_cuda and taps.is_contiguous() and taps.dtype == torch.float32 and taps.numel() == k * k
    assert active.dtype == torch.int32 and active.numel() == B and active.is_cuda
    _launch("blur2d", lambda: load().mphsir_blur2d(x.data_ptr(), out.data_ptr(), taps.data_ptr(), active.data_ptr(), B, Cc, H, W, k,
                                                   stream_ptr()),
            lambda: (2.0 * k * k * x.numel(), 8.0 * x.numel(), "blur2d"))


def topk_mean(x: torch.Tensor, k: int) -> torch.Tensor:
    """[B,C,H,W] -> [B,C]: mean of the k largest pixels of every plane (utils/degradation_utils.py:257-261)"""
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    mean = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    _launch("topk_mean", lambda: load().mphsir_topk_mean(x.data_ptr(), B * Cc, H * W, k, mean.data_ptr(), stream_ptr()),
            lambda: (0.0, 4.0 * x.numel() * k, "topk_mean"))
    return mean


def haze(x: torch.Tensor, out: torch.Tensor, cirrus: torch.Tensor, omega: torch.Tensor, expo: torch.Tensor,
         light: torch.Tensor) -> None:
    """out[b] = x[b] * T + light[b,c] * (1 - T) where omega[b] > 0 (utils/degradation_utils.py:263-271)"""
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape and x.dtype == torch.float32
    for t, n in ((cirrus, B * H * W), (omega, B), (expo, Cc), (light, B * Cc)):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == n
    _launch("haze", lambda: load().mphsir_haze(x.data_ptr(), out.data_ptr(), cirrus.data_ptr(), omega.data_ptr(), expo.data_ptr(),
                                               light.data_ptr(), B, Cc, H * W, stream_ptr()),
            lambda: (0.0, 12.0 * x.numel(), "haze"))


def poisson(x: torch.Tensor, out: torch.Tensor, scale: torch.Tensor, seed: int) -> None:
    """out[b] = Poisson(max(x[b], 0) * scale[b]) / scale[b] where scale[b] > 0 (utils/degradation_utils.py:86-89); Philox stream
    keyed by `seed`, counter word 2 = 2."""
    B = x.shape[0]
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape and x.dtype == torch.float32
    assert scale.is_cuda and scale.dtype == torch.float32 and scale.numel() == B
    _launch("poisson", lambda: load().mphsir_poisson(x.data_ptr(), out.data_ptr(), scale.data_ptr(), B, x.numel() // B,
                                                     seed & 0xFFFFFFFFFFFFFFFF, stream_ptr()),
            lambda: (0.0, 8.0 * x.numel(), "poisson"))


def degrade(clean: torch.Tensor, out: torch.Tensor, sigma: torch.Tensor, keep: torch.Tensor, mask_ratio: torch.Tensor,
            seed: int) -> None:
    """out = clean * keep[b,c] * (u > mask_ratio[b]) + sigma[b,c] * n  (Philox4x32-10 stream keyed by `seed`)"""
    B, Cc, H, W = clean.shape
    assert clean.is_contiguous() and out.is_contiguous() and out.shape == clean.shape and clean.dtype == torch.float32
    assert sigma.numel() == B * Cc and keep.numel() == B * Cc and mask_ratio.numel() == B
    _launch("degrade", lambda: load().mphsir_degrade(clean.data_ptr(), out.data_ptr(), B, Cc, H * W, sigma.data_ptr(), keep.data_ptr(),
                                                     mask_ratio.data_ptr(), seed & 0xFFFFFFFFFFFFFFFF, stream_ptr()),
            lambda: (0.0, 8.0 * clean.numel(), "degrade"))
