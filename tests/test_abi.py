"""CPU: the C-ABI library builds, loads, and exports every symbol include/mphsir.h declares.
No compute call is made (there is no GPU in the authoring container)."""
import ctypes
import os
import re
import subprocess

import pytest

from mp_hsir_b200 import build, lib
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def so_path():
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mphsir.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mphsir_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    assert "mphsir_gemm_fwd" in syms and "mphsir_window_attn_fwd" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(so_path):
    dll = ctypes.CDLL(so_path)
    for s in declared_symbols():
        assert hasattr(dll, s), f"{s} declared in mphsir.h but not exported"
    assert dll.mphsir_version() == 100


def test_python_binding_covers_header(so_path):
    assert sorted(lib.SIGNATURES) == declared_symbols()
    lib.load()


def test_only_abi_symbols_are_exported(so_path):
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("mphsir_") for s in exported), exported


def test_library_is_sm100a_only(so_path):
    out = subprocess.run(["cuobjdump", "-lelf", so_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header():
    # 64-bit pointers, natural alignment: sizes computed by hand from include/mphsir.h
    # compile a C probe against the real header and compare sizeof / offsetof with the ctypes mirrors
    import subprocess, tempfile, os
    src = r"""
#include <stdio.h>
#include <stddef.h>
#include "mphsir.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu ", sizeof(mphsir_wgrad_params), offsetof(mphsir_wgrad_params, dw_batch_stride),
         offsetof(mphsir_wgrad_params, so), offsetof(mphsir_wgrad_params, precision), offsetof(mphsir_wgrad_params, dbias),
         sizeof(mphsir_local_gate_bwd_weights));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mphsir_gemm_params), offsetof(mphsir_gemm_params, row_scale),
         offsetof(mphsir_gemm_params, Bimg), offsetof(mphsir_gemm_params, bimg_batch_bytes),
         offsetof(mphsir_gemm_params, Y2), offsetof(mphsir_gemm_params, n_split),
         sizeof(mphsir_conv3x3_params), offsetof(mphsir_conv3x3_params, Bimg), sizeof(mphsir_local_gate_params));
  return 0;
}
"""
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    G, Cv, L = lib.GemmParams, lib.ConvParams, lib.LocalGateParams
    Wg = lib.WgradParams
    want = [ctypes.sizeof(Wg), Wg.dw_batch_stride.offset, Wg.so.offset, Wg.precision.offset, Wg.dbias.offset,
            ctypes.sizeof(lib.LocalGateBwdWeights)]
    want += [ctypes.sizeof(G), G.row_scale.offset, G.Bimg.offset, G.bimg_batch_bytes.offset, G.Y2.offset,
            G.n_split.offset, ctypes.sizeof(Cv),
            Cv.Bimg.offset, ctypes.sizeof(L)]
    assert got == want


def test_integration_md_struct_is_the_current_layout():
    """INTEGRATION.md shows maintainers a ctypes mirror of mphsir_gemm_params: it must be the struct the library reads
    (round 1 shipped a stale one)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class GemmParams\(ctypes.Structure\):.*?\n    _fields_ = \[(.*?)\]\n\n", text, flags=re.S)
    assert m, "GemmParams snippet not found in INTEGRATION.md"
    ns = {"ctypes": ctypes}
    exec("class GemmParams(ctypes.Structure):\n    _fields_ = [" + m.group(1) + "]", ns)
    doc = ns["GemmParams"]
    ours = lib.GemmParams
    assert [(n, t) for n, t in doc._fields_] == [(n, t) for n, t in ours._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(ours)
