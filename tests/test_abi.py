"""CPU: the C-ABI library loads and exports every symbol include/swb200.h declares; without a GPU
the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from simpleworks_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_and_library_agree(lib):
    from simpleworks_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "swb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(swb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_arkworks_layout():
    # Fp256 = 32 B, Fp384 = 48 B, GroupAffine = 2*48 + bool (padded to 104), Jacobian = 144 B
    hdr = open(os.path.join(ROOT, "include", "swb200.h")).read()
    assert "uint64_t l[4]; } swb_fr" in hdr and "uint64_t l[6]; } swb_fq" in hdr
    assert "swb_fq x, y; uint8_t infinity; uint8_t _pad[7]; } swb_g1_affine" in hdr


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.swb_init(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.swb_last_error(None)
    from simpleworks_b200.binding import Backend, SwbError
    with pytest.raises(SwbError):
        Backend(0)


def test_product_never_imports_the_oracle():
    """No file of the product includes, imports, links or calls anything under oracle/ (comments
    may say that the oracle build re-uses the protocol templates; the dependency is one-way)."""
    pkg = os.path.join(ROOT, "simpleworks_b200")
    bad = re.compile(r'#include\s*[<"][^>"]*oracle|\bimport\s+oracle|\bfrom\s+oracle|liboracle|\borc_[a-z0-9_]+\s*\(|pyoracle|pymarlin')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert not bad.search(src), os.path.join(root, f)
