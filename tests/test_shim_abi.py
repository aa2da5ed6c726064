"""CPU tier: the three statements of the C ABI must agree -- include/g16b200.h (the contract), groth16_b200/_lib.py (the
ctypes binding every test uses) and shim/ark-groth16-b200/src/sys.rs (the Rust binding a maintainer links against; source
only here, no Rust toolchain in this image).  Checked: the set of functions, the arity of each, pointer-ness of every
parameter, and the field lists of the structs that cross the boundary."""
import os
import re

from groth16_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_comments(txt):
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    return re.sub(r"//[^\n]*", " ", txt)


def header_functions():
    txt = _strip_comments(open(os.path.join(ROOT, "include", "g16b200.h")).read())
    out = {}
    for m in re.finditer(r"\b(?:int|void|uint32_t|const char\s*\*)\s*(g16_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        out[name] = ["*" in p for p in params]
    return out


def header_structs():
    txt = _strip_comments(open(os.path.join(ROOT, "include", "g16b200.h")).read())
    out = {}
    for m in re.finditer(r"typedef struct\s*\{(.*?)\}\s*(g16_\w+)\s*;", txt, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for piece in decl.split(","):
                nm = re.findall(r"(\w+)\s*(?:\[\s*\d+\s*\])?\s*$", piece.strip())
                if nm:
                    fields.append(nm[0])
        out[m.group(2)] = fields
    return out


def rust_functions():
    txt = _strip_comments(open(os.path.join(ROOT, "shim", "ark-groth16-b200", "src", "sys.rs")).read())
    block = re.search(r'extern "C"\s*\{(.*)\}', txt, flags=re.S).group(1)
    out = {}
    for m in re.finditer(r"pub fn (g16_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        args = m.group(2).strip()
        params = [a.strip() for a in args.split(",") if a.strip()]
        out[m.group(1)] = ["*" in p.split(":", 1)[1] for p in params]
    return out


def rust_structs():
    txt = _strip_comments(open(os.path.join(ROOT, "shim", "ark-groth16-b200", "src", "sys.rs")).read())
    out = {}
    for m in re.finditer(r"pub struct (g16_\w+)\s*\{(.*?)\}", txt, flags=re.S):
        out[m.group(1)] = re.findall(r"pub (\w+)\s*:", m.group(2))
    return out


def test_header_ctypes_and_rust_declare_the_same_functions():
    h, r = header_functions(), rust_functions()
    py = {name: args for name, _, args in _lib.SIGNATURES}
    assert len(h) >= 26
    assert set(h) == set(py), sorted(set(h) ^ set(py))
    assert set(h) == set(r), sorted(set(h) ^ set(r))
    for name, ptrs in h.items():
        assert len(py[name]) == len(ptrs), name
        assert r[name] == ptrs, (name, r[name], ptrs)


def test_struct_layouts_agree():
    h, r = header_structs(), rust_structs()
    for name in ("g16_csr", "g16_pk_desc", "g16_pk_export_desc", "g16_timings", "g16_config"):
        assert h[name] == r[name], (name, h[name], r[name])
    assert [f for f, _ in _lib.Csr._fields_] == h["g16_csr"]
    assert [f for f, _ in _lib.PkDesc._fields_] == h["g16_pk_desc"]
    assert [f for f, _ in _lib.PkExportDesc._fields_] == h["g16_pk_export_desc"]
    assert [f for f, _ in _lib.Timings._fields_] == h["g16_timings"]
    assert [f for f, _ in _lib.Config._fields_] == h["g16_config"]


def test_shim_source_is_complete():
    """every helper the shim's public functions call is written out (round 1 shipped a sketch with `unimplemented!()`)"""
    src = open(os.path.join(ROOT, "shim", "ark-groth16-b200", "src", "lib.rs")).read()
    assert "unimplemented!" not in src and "todo!" not in src and "/* ... */" not in src
    for item in ("fn pack_points", "fn unpack_point", "impl Csr", "fn with_thread_ctx", "fn load_matrices_once",
                 "impl R1CSToQAP for GpuReduction", "impl<E: SwPairing> SNARK<E::ScalarField> for Groth16B200<E>",
                 "fn create_proof_with_reduction_and_matrices", "fn create_random_proof_with_reduction"):
        assert item in src, item
    body = re.search(r'extern "C"\s*\{(.*)\}', open(os.path.join(ROOT, "shim", "ark-groth16-b200", "src", "sys.rs")).read(), flags=re.S).group(1)
    for call in set(re.findall(r"sys::(g16_\w+)\s*\(", src)):
        assert f"fn {call}" in body, call
