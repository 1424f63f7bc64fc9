"""Architecture description shared by the CUDA module, the oracle and the tests.

Mirrors the constructor arguments of the reference (net/MP_HSIR.py:764-775) and the
per-stage hyper-parameters hard-coded at net/MP_HSIR.py:791-805.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

WINDOW = 8          # window_size=[8,8,8]            (net/MP_HSIR.py:769)
SHIFT = 4           # window_size // 2 on odd blocks (net/MP_HSIR.py:748)
PROMPT_LEN = 128    # spectral prompt length          (net/MP_HSIR.py:791-805)
LN_EPS = 1e-5
CLIP_DIM = 512


@dataclass(frozen=True)
class Stage:
    name: str
    depth: int
    dim: int
    heads: int
    compress: int            # compress_ratio -> low rank r = dim // compress
    construct_res: int       # input_resolution baked into the attn_mask buffer
    dpr: Tuple[float, ...]   # DropPath rate per block (train mode only)

    @property
    def rank(self) -> int:
        return self.dim // self.compress


@dataclass(frozen=True)
class NetConfig:
    in_channel: int = 31
    out_channel: int = 31
    dim: int = 64
    num_blocks: Tuple[int, int, int] = (2, 4, 6)
    window_size: Tuple[int, int, int] = (8, 8, 8)
    task_classes: int = 6
    num_refinement_blocks: int = 4
    heads: Tuple[int, int, int] = (2, 4, 8)
    ffn_expansion_factor: float = 2.66
    bias: bool = False

    def __post_init__(self):
        if self.task_classes not in (1, 6, 7):
            # same failure as Text_Prompt.__init__ (net/MP_HSIR.py:507-508)
            raise ValueError("task_classes must be 6 or 7")
        if tuple(self.window_size) != (8, 8, 8):
            raise ValueError("only the reference window_size=[8,8,8] is supported by the CUDA path")
        if self.bias:
            raise ValueError("bias=True is never used by the reference call sites (train.py:45, test.py:39)")

    @staticmethod
    def natural() -> "NetConfig":
        return NetConfig(31, 31, 64, task_classes=6)      # test.py:39

    @staticmethod
    def remote_sensing() -> "NetConfig":
        return NetConfig(100, 100, 96, task_classes=7)    # train.py:45

    def hidden(self, dim: int) -> int:
        return int(dim * self.ffn_expansion_factor)       # net/MP_HSIR.py:621, :378

    def drop_path_rates(self) -> List[float]:
        n = sum(self.num_blocks)                          # net/MP_HSIR.py:780
        if n == 1:
            return [0.0]
        return [0.1 * i / (n - 1) for i in range(n)]

    def stages(self) -> List[Stage]:
        d, nb, hd = self.dim, self.num_blocks, self.heads
        dpr = self.drop_path_rates()
        s0, s1, s2 = 0, nb[0], nb[0] + nb[1]
        d0 = tuple(dpr[s0:s1]); d1 = tuple(dpr[s1:s2]); d2 = tuple(dpr[s2:s2 + nb[2]])
        ref = tuple((list(d1) + [0.0] * self.num_refinement_blocks)[: self.num_refinement_blocks])
        return [
            Stage("encoder_level1", nb[0], d, hd[0], 8, 64, d0),
            Stage("encoder_level2", nb[1], 2 * d, hd[1], 16, 32, d1),
            Stage("latent", nb[2], 4 * d, hd[2], 32, 16, d2),
            Stage("decoder_level2", nb[1], 2 * d, hd[1], 16, 32, d1),
            Stage("decoder_level1", nb[0], 2 * d, hd[0], 8, 64, d0),
            Stage("refinement", self.num_refinement_blocks, 2 * d, hd[0], 8, 64, ref),
        ]
