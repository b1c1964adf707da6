"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/nvit_b200.h declares, with matching arity in the ctypes binding.  No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nvit_b200.h")
TUNING_HEADER = os.path.join(ROOT, "include", "nvit_b200_tuning.h")


def declared_functions(path=HEADER, with_hooks=False):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    if not with_hooks:       # the -DNVIT_BENCH_HOOKS section is not part of the product library
        text = re.sub(r"#ifdef NVIT_BENCH_HOOKS.*?#endif", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(nvit_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib():
    from nvit_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_the_hot_path_entry_points():
    fns = declared_functions()
    for name in ("nvit_gemm_bf16", "nvit_residual_fwd", "nvit_residual_bwd", "nvit_attention_fwd", "nvit_attention_bwd",
                 "nvit_weight_norm_multi", "nvit_im2col_bf16", "nvit_adamw_flat", "nvit_cross_entropy"):
        assert name in fns


def test_library_exports_every_declared_symbol(lib):
    from nvit_b200 import _lib
    fns = declared_functions()
    assert len(fns) >= 20
    for name, nargs in fns.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, binding has {len(_lib.SIGNATURES[name])}"
    tuning = declared_functions(TUNING_HEADER)
    for name, nargs in tuning.items():
        assert hasattr(lib, name), f"{name} declared in the tuning header but not exported"
        assert len(_lib.SIGNATURES[name]) == nargs
    for name in _lib.SIGNATURES:
        assert name in fns or name in tuning, f"{name} bound but not declared in a header"
    # wrong-output measurement hooks are compiled out of the product library
    hooks = set(declared_functions(TUNING_HEADER, with_hooks=True)) - set(tuning)
    assert hooks == set(_lib.HOOK_SIGNATURES) and hooks
    for name in hooks:
        assert not hasattr(lib, name), f"{name} must exist only in -DNVIT_BENCH_HOOKS builds"
    # and the boundary header itself declares no process-wide variant switches
    for name in tuning:
        assert name not in fns


def test_library_calls_without_gpu(lib):
    assert lib.nvit_version() >= 100
    assert lib.nvit_sm_count() > 0


def test_argument_errors_are_reported_not_thrown(lib):
    from nvit_b200 import _lib
    with pytest.raises(RuntimeError, match="null operand"):
        _lib.call("nvit_gemm_bf16", None, None, None, None, 1, 1, 1, 8, 8, 8, 0, 0, 0, 0, 0, 1, None, None, 1.0, None, 0, 0, None)
    with pytest.raises(RuntimeError, match="head_dim must be 64"):
        _lib.call("nvit_attention_fwd", 16, 16, 16, 8, 8, 8, None, 1.0, 1.0, 16, 8, 16, 1, 1, 4, 32, None, None, 0, 0, None)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        _lib.call("nvit_residual_fwd", 16, 16, 16, 1.0, None, None, 16, None, 4, 7, None)


def test_missing_library_fails_loudly(monkeypatch):
    from nvit_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnvit_b200.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.load()


def test_sm_budget_window_counts_entry_point_calls():
    """Data-parallel overlap: after a bucket's all-reduce the next n entry-point calls size their grids for fewer SMs, then all
    SMs again (nvit_b200._lib.sm_budget_window; no kernel runs here)."""
    from nvit_b200 import _lib
    lib = _lib.load()
    total = lib.nvit_sm_count()
    try:
        _lib.sm_budget_window(3, total - 8)
        assert lib.nvit_sm_count() == total - 8
        _lib.call("nvit_set_pdl", 0)
        _lib.call("nvit_set_pdl", 0)
        assert lib.nvit_sm_count() == total - 8
        _lib.call("nvit_set_pdl", 0)
        assert lib.nvit_sm_count() == total
        _lib.call("nvit_set_pdl", 0)
        assert lib.nvit_sm_count() == total
    finally:
        lib.nvit_set_sm_budget(0)
