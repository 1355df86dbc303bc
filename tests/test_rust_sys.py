"""rust/lightgarden-cuda-sys/src/lib.rs is held to include/light_garden_b200.h: every entry point, every argument,
every struct field (order and type), every constant -- and to light_garden_b200/abi.py, the ctypes twin the tests use.
Round 1 shipped the crate as markdown with 23 of 43 entry points; now it is generated (tools/gen_rust_sys.py) and this
test fails on any drift.  The patches of rust/patches/ are applied to a scratch copy of the reference when it is there."""
import ctypes as C
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_rust_sys  # noqa: E402

LIB_RS = os.path.join(ROOT, "rust", "lightgarden-cuda-sys", "src", "lib.rs")
RUST_SIZE = {"i32": 4, "u32": 4, "f32": 4, "i64": 8, "u64": 8, "f64": 8, "usize": 8}


def parse_lib_rs():
    text = open(LIB_RS).read()
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (\w+): i32 = (-?\d+);", text)}
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\]\n(?:#\[derive\([^)]*\)\]\n)?pub struct (\w+) \{(.*?)\n\}", text, flags=re.S):
        fields = re.findall(r"pub (\w+): ([^,\n]+),", m.group(2))
        structs[m.group(1)] = fields
    fns = {}
    ext = text[text.index('extern "C" {'):]
    for m in re.finditer(r"pub fn (\w+)\((.*?)\) -> ([^;]+);", ext):
        args = [a.strip().split(": ", 1) for a in m.group(2).split(", ") if a.strip()]
        fns[m.group(1)] = (args, m.group(3).strip())
    return consts, structs, fns


def test_lib_rs_is_what_the_header_generates():
    assert open(LIB_RS).read() == gen_rust_sys.generate(), "run python tools/gen_rust_sys.py"


def test_every_entry_point_argument_and_field_matches_the_header():
    defines, enums, hstructs, hfns = gen_rust_sys.parse_header()
    consts, structs, fns = parse_lib_rs()
    assert consts == dict(defines + enums)
    assert set(fns) == {f[0] for f in hfns} and len(hfns) >= 46
    for name, ret, args in hfns:
        rargs, rret = fns[name]
        assert rret == gen_rust_sys.rust_type(ret), name
        assert [t for _, t in rargs] == [gen_rust_sys.rust_type(ct) for _, ct in args], name
        assert [n.rstrip("_") for n, _ in rargs] == [n for n, _ in args], name
    for sname, fields in hstructs:
        assert [f for f, _ in structs[sname]] == [f for f, _, _ in fields], sname


def _rust_sizeof(structs, ty):
    m = re.match(r"\[(\w+); (\d+)\]", ty)
    if m:
        return _rust_sizeof(structs, m.group(1)) * int(m.group(2))
    if ty in RUST_SIZE:
        return RUST_SIZE[ty]
    return sum(_rust_sizeof(structs, t) for _, t in structs[ty])     # no padding: checked against the C sizes below


def test_struct_layouts_agree_with_the_ctypes_binding():
    """Same field names in the same order, and the sizes the header states (static_asserts in csrc/lg_capi.cu)."""
    from light_garden_b200 import abi
    _, structs, fns = parse_lib_rs()
    for name, fields in structs.items():
        if name == "lg_ctx":
            continue
        ct = getattr(abi, name, None)
        if ct is not None and hasattr(ct, "_fields_"):
            assert [f for f, _ in fields] == [f[0] for f in ct._fields_], name
            assert _rust_sizeof(structs, name) == C.sizeof(ct), name
        if name in abi.SIZES:
            assert _rust_sizeof(structs, name) == abi.SIZES[name][1], name
    assert set(fns) == set(abi.PROTOTYPES), set(fns) ^ set(abi.PROTOTYPES)


def test_the_shim_uses_only_entry_points_and_fields_that_exist():
    consts, structs, fns = parse_lib_rs()
    shim = open(os.path.join(ROOT, "rust", "patches", "cuda.rs")).read()
    for name in set(re.findall(r"cu::(lg_\w+)", shim)):
        assert name in fns or name == "lg_ctx", name
    for name in set(re.findall(r"cu::(LG_\w+)", shim)):
        assert name in consts, name
    for sname, body in re.findall(r"(?<!-> )cu::(Lg\w+) \{([^}]*)\}", shim):
        named = set(re.findall(r"(?<![:\w.])([a-z_]\w*):(?!:)", body))                 # `field: value`
        shorthand = {w for w in re.findall(r"(?:^|,)\s*([a-z_]\w*)\s*(?=,|$)", body)}     # `field,`
        fields = {f for f, _ in structs[sname]}
        assert named <= fields, (sname, named - fields)
        assert fields <= named | shorthand, (sname, fields - named - shorthand)           # a struct literal names them all


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present (GPU box)")
def test_patches_apply_to_the_reference(tmp_path):
    work = tmp_path / "light_garden"
    shutil.copytree(REF, work, ignore=shutil.ignore_patterns(".git", "images", "target", "*.jpg"))
    subprocess.run(["git", "init", "-q"], cwd=work, check=True)
    for p in sorted(os.listdir(os.path.join(ROOT, "rust", "patches"))):
        if p.endswith(".patch"):
            r = subprocess.run(["git", "apply", "--check", "-p1", os.path.join(ROOT, "rust", "patches", p)], cwd=work,
                               capture_output=True, text=True)
            assert r.returncode == 0, (p, r.stderr)
            subprocess.run(["git", "apply", "-p1", os.path.join(ROOT, "rust", "patches", p)], cwd=work, check=True)
    text = (work / "src" / "light_garden" / "tracer.rs").read_text()
    assert "cu.sync(" in text and "pub cuda: Option<cuda::CudaPath>" in text
