"""CPU tests of the drop-in boundary: libbevr.so loads and exports every symbol of include/bevr.h;
argument validation works without a GPU (no kernel is launched)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bevr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bevr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = header_symbols()
    for s in ("bevr_msda_fwd", "bevr_msda_bwd", "bevr_project_points", "bevr_rpe_attn_fwd",
              "bevr_rpe_attn_bwd", "bevr_last_error", "bevr_version"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from bevrender_b200 import _lib
    handle = C.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else C.CDLL(_lib.build())
    for s in header_symbols():
        assert hasattr(handle, s), f"libbevr.so lacks {s}"
    assert set(header_symbols()) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib().bevr_version() == _lib.EXPECTED_VERSION == 210


def test_argument_validation_returns_negative_and_message():
    from bevrender_b200 import _lib
    L = _lib.lib()
    hw = (C.c_int32 * 2)(4, 4)
    st = (C.c_int32 * 1)(0)
    # Dh=6 fp32 is not a multiple of 16 bytes -> rejected before any CUDA call
    rc = L.bevr_msda_fwd(None, hw, st, None, None, None, None, 1, 16, 1, 6, 4, 1, 1, 0, 0, 0, None)
    assert rc < 0 and b"16 bytes" in L.bevr_last_error()
    rc = L.bevr_msda_fwd(None, hw, st, None, None, None, None, 1, 8, 1, 4, 4, 1, 1, 0, 0, 0, None)
    assert rc < 0 and b"outside S" in L.bevr_last_error()
    rc = L.bevr_project_points(None, None, None, None, None, None, 4, 160, 128, None)
    assert rc < 0


def test_empty_inputs_are_a_successful_no_op():
    """Empty batch / empty query list: the entry points return 0 without touching the device."""
    from bevrender_b200 import _lib
    L = _lib.lib()
    hw = (C.c_int32 * 2)(4, 4)
    st = (C.c_int32 * 1)(0)
    assert L.bevr_msda_fwd(None, hw, st, None, None, None, None, 0, 16, 1, 4, 8, 1, 1, 0, 0, 0, None) == 0
    assert L.bevr_msda_fwd(None, hw, st, None, None, None, None, 2, 16, 1, 4, 0, 1, 1, 0, 0, 0, None) == 0
    assert L.bevr_msda_bwd(None, hw, st, None, None, None, None, None, None, None, None, 0, 16, 1, 4, 8, 1, 1, 0, 0, 0,
                           None) == 0
    assert L.bevr_project_points(None, (C.c_float * 16)(), (C.c_float * 9)(), None, None, None, 0, 160, 128, None) < 0
    assert L.bevr_msda_fwd(None, hw, st, None, None, None, None, -1, 16, 1, 4, 8, 1, 1, 0, 0, 0, None) < 0


def test_operators_refuse_cpu_tensors():
    import torch
    import bevrender_b200 as bb
    v = torch.zeros(1, 16, 1, 4)
    loc = torch.zeros(1, 4, 1, 1, 1, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        bb.ms_deform_attn(v, [(4, 4)], [0], loc)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bevrender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_offset_head_validation_and_empty_batch():
    from bevrender_b200 import _lib
    L = _lib.lib()
    # 9 output channels exceed the kernel's 8; the plain layout needs exactly 2; the interleaved one an even H
    assert L.bevr_offset_head_fwd(None, None, None, None, None, 0, None, None, 1, 8, 4, 4, 9, 1, 1, 0.5, 1e-5, None) < 0
    assert b"output channels" in L.bevr_last_error()
    assert L.bevr_offset_head_fwd(None, None, None, None, None, 0, None, None, 1, 8, 4, 4, 3, 0, 1, 0.5, 1e-5, None) < 0
    assert L.bevr_offset_head_fwd(None, None, None, None, None, 0, None, None, 1, 8, 5, 4, 4, 1, 1, 0.5, 1e-5, None) < 0
    assert L.bevr_offset_head_fwd(None, None, None, None, None, 0, None, None, 1, 8, 4, 4, 2, 0, 1, 0.5, 1e-5, None) < 0  # NULLs
    assert L.bevr_offset_head_fwd(None, None, None, None, None, 0, None, None, 0, 8, 4, 4, 2, 0, 1, 0.5, 1e-5, None) == 0
    assert L.bevr_offset_head_bwd(None, None, None, None, None, 0, None, None, None, None, None, None, 0, 8, 4, 4, 2, 0, 1,
                                  0.5, 1e-5, None) == 0


def test_rpe_attention_workspace_covers_the_forward_split():
    """The forward's key-range partials and the pre-split tf32 operands live in the caller's workspace: the size query
    must cover the padded table, the partials when the grid is small enough to be split, and 4 operand copies."""
    from bevrender_b200 import _lib
    L = _lib.lib()
    table = 2 * (111 + 4) * (111 + 4) * 4
    kv = 4 * 16 * 3136 * 32 * 4                            # K/V hi/lo, Ns = 3136 is a multiple of 64
    small = L.bevr_rpe_attn_workspace_bytes(16, 2, 32, 3136, 3136, 111, 111, 0)
    assert small > table + kv + 16 * 34 * 3136 * 4         # at least two ranges of (Dh + 2) floats per row
    big = L.bevr_rpe_attn_workspace_bytes(8, 8, 32, 40000, 40000, 399, 399, 0)
    assert big == 8 * 403 * 403 * 4 + 4 * 8 * 40000 * 32 * 4 + 256      # 5000 CTAs already: no split
    bwd = L.bevr_rpe_attn_workspace_bytes(8, 8, 32, 40000, 40000, 399, 399, 1)
    assert bwd >= 8 * 403 * 403 * 4 * 2 + 8 * 8 * 40000 * 32 * 4       # + Q / dO copies and >= 1 gradient table
