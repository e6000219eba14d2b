"""CPU-side checks of the drop-in boundary: the shared library loads and exports exactly the
symbols include/ct_b200.h declares; the ctypes mirror covers them; the product path refuses to run
without CUDA (no CPU fallback). No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ct_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(ct_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from cleantransformer_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_lib.SIGNATURES) == syms, (set(syms) ^ set(_lib.SIGNATURES))
    assert _lib.load().ct_version() == 100


def test_struct_mirrors_match_header_sizes(tmp_path):
    """sizeof/offsetof of the argument structs as gcc sees include/ct_b200.h == the ctypes mirrors."""
    import subprocess
    from cleantransformer_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ct_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ct_gemm_args), sizeof(ct_attn_args), '
        'sizeof(ct_attn_bwd_args), offsetof(ct_gemm_args, residual), offsetof(ct_attn_args, kbias2), '
        'offsetof(ct_attn_bwd_args, delta), sizeof(ct_ln_bwd_args), offsetof(ct_ln_bwd_args, dxsum));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.GemmArgs), ctypes.sizeof(_lib.AttnArgs), ctypes.sizeof(_lib.AttnBwdArgs),
            _lib.GemmArgs.residual.offset, _lib.AttnArgs.kbias2.offset, _lib.AttnBwdArgs.delta.offset,
            ctypes.sizeof(_lib.LnBwdArgs), _lib.LnBwdArgs.dxsum.offset]
    assert got == want


def test_error_convention_and_no_cpu_fallback():
    from cleantransformer_b200 import _lib, ops
    lib = _lib.load()
    rc = lib.ct_device_check(0)
    if not torch.cuda.is_available():
        assert rc == -2  # CT_ERR_UNSUPPORTED: no device
        assert "CUDA" in _lib.last_error() or "device" in _lib.last_error()
    # bad arguments are rejected before any launch
    assert lib.ct_gemm(None, None) == -1
    assert "null" in _lib.last_error()
    assert lib.ct_layernorm_fwd(None, 0, None, None, None, 0, None, 0, None, None, 1, 1, 1e-5, None) == -1
    assert lib.ct_allreduce_bucket(0, 16, 1.0, 0, 0, None) == -4  # comm not initialised
    # the fused gather + LayerNorm entry point: shape / pointer / table-order / alignment checks precede any launch
    f = lib.ct_embedding_layernorm_fwd
    a = 1 << 12  # any non-null, 16-byte aligned address: it is never dereferenced on these paths
    assert f(a, a, 10, None, None, 0, None, None, 0, a, a, None, a, 0, None, 0, None, None, -1, 128, 1e-5, None) == -1
    assert f(None, None, 10, None, None, 0, None, None, 0, a, a, None, a, 0, None, 0, None, None, 4, 128, 1e-5, None) == -1
    assert f(a, a, 10, None, None, 0, a, a, 5, a, a, None, a, 0, None, 0, None, None, 4, 128, 1e-5, None) == -1
    assert "in order" in _lib.last_error()
    assert f(a, a, 10, None, None, 0, None, None, 0, a, a, None, a, 0, None, 0, None, None, 4, 96, 1e-5, None) == -2
    assert "ct_embedding_fwd + ct_layernorm_fwd" in _lib.last_error()
    assert f(a, a + 4, 10, None, None, 0, None, None, 0, a, a, None, a, 0, None, 0, None, None, 4, 128, 1e-5, None) == -2
    assert f(a, a, 10, None, None, 0, None, None, 0, a, a, None, a, 7, None, 0, None, None, 4, 128, 1e-5, None) == -2
    assert f(a, a, 10, None, None, 0, None, None, 0, a, a, None, a, 0, None, 0, None, None, 0, 128, 1e-5, None) == 0  # empty
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.layernorm_fwd(x, torch.ones(8), torch.zeros(8), 1e-5)
    from cleantransformer_b200.optimizer import AdamW
    p = torch.nn.Parameter(torch.randn(3))
    p.grad = torch.randn(3)
    with pytest.raises(RuntimeError, match="CUDA"):
        AdamW([p]).step()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the product may import it."""
    pkg = os.path.join(ROOT, "cleantransformer_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), os.path.join(dirpath, f)
                assert "ct_oracle" not in src, os.path.join(dirpath, f)
