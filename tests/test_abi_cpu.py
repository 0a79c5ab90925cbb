"""CPU suite: the C-ABI library builds, loads and exports every symbol include/branchedgp_b200.h declares."""
import ctypes
import os
import re

from branchedgp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "branchedgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert "bgp_dense_elbo_grad" in names and "bgp_create" in names
    lib = ctypes.CDLL(_lib.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in Python but not declared in the header"


def test_pure_queries_work_without_a_gpu():
    lib = _lib.load_library()
    assert lib.bgp_version() >= 100
    assert lib.bgp_debug_padded_size(1000) == 3072
    assert lib.bgp_debug_padded_size(20) == 128
    assert lib.bgp_dense_slot_bytes(1000, 1) > 4 * (3000 * 3001 // 2) * 8


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "branchedgp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r"#.*", "", txt).replace("no oracle", ""), f"{f} mentions the oracle"
