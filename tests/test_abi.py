"""The C-ABI library loads and exports every symbol include/*.h declares; struct layouts match.
No compute calls: runs without a GPU."""
import ctypes
import os
import re

from light_garden_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "light_garden_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", text)))


def test_struct_sizes_match_header_comments():
    for name, (got, want) in abi.SIZES.items():
        assert got == want, name


def test_every_declared_symbol_is_exported_and_bound(product_lib):
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(product_lib, n), f"{n} declared in the header but not exported"
        assert n in abi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(abi.PROTOTYPES) == names


def test_abi_version(product_lib):
    assert product_lib.lg_abi_version() == abi.LG_ABI_VERSION


def test_create_fails_loudly_without_device(product_lib):
    """No CPU fallback: without a CUDA device lg_create must fail, not limp along."""
    n = ctypes.c_int32(-1)
    rc = product_lib.lg_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        return  # GPU box: covered by the gpu tests
    h = ctypes.c_void_p()
    assert product_lib.lg_create(0, abi.LG_PRECISION_F32, ctypes.byref(h)) == abi.LG_ERR_CUDA
    assert not h.value


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under light_garden_b200/ (nor the tools/ around it) may import,
    include or load it."""
    walk = list(os.walk(os.path.join(ROOT, "light_garden_b200"))) + list(os.walk(os.path.join(ROOT, "tools")))
    for dirpath, _, files in walk:
        if "_lib" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(dirpath, f)).read()
            bad = re.search(r'(^|\n)\s*(import\s+(lg_)?oracle|from\s+(lg_)?oracle|from\s+\.+\s*import\s+oracle)'
                            r'|liblg_oracle|lgo_[a-z_]+\s*\(|#include\s+"[^"]*oracle/', text)
            assert not bad, f"{f} references the oracle: {bad.group(0)!r}"
