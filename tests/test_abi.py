"""CPU checks of the drop-in boundary: the shared library loads, exports exactly the symbols include/anemoi_b200.h
declares, and the ctypes table covers them (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

from anemoi_core_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "anemoi_b200.h")).read()
    return set(re.findall(r"ANEMOI_API\s+[\w\s\*]+?\b(anemoi_b200_\w+)\s*\(", src))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    assert {"anemoi_b200_csr_build", "anemoi_b200_layer_norm", "anemoi_b200_linear", "anemoi_b200_gt_attention_fwd",
            "anemoi_b200_graphconv_ln_aggregate", "anemoi_b200_cast_pad", "anemoi_b200_add", "anemoi_b200_row_stats", "anemoi_b200_last_error",
            "anemoi_b200_abi_version"} <= syms  # fmt: skip


def test_library_exports_every_header_symbol_and_ctypes_table_matches():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == syms
    # argument counts of the ctypes table == parameter counts in the header
    src = open(os.path.join(ROOT, "include", "anemoi_b200.h")).read()
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"ANEMOI_API[^;(]*\b" + name + r"\s*\(([^;]*?)\)\s*;", src, re.S)
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else params.count(",") + 1
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes table {len(argtypes)}"


def test_version_and_error_string_calls_work_without_a_gpu():
    lib = _lib.load()
    assert lib.anemoi_b200_abi_version() == 6
    assert isinstance(lib.anemoi_b200_last_error(), bytes)
    # argument validation happens before any CUDA call: a bad shape is reported through the error channel
    rc = lib.anemoi_b200_linear(None, 0, None, 0, 0, None, None, None, None, None, 0, None, 0, 0, None, 0, 0, -1, 1, 1, 0, None, None, 0, 0, 0.0, None, None)
    assert rc == -1 and b"bad shape" in lib.anemoi_b200_last_error()
