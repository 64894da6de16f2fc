"""Static check of the drop-in claim: every `pub fn` of the reference's `impl BinaryQuantizer`
(src/quantization.rs) and every method of `trait VectorIndex` (src/index.rs) must exist, with the same
number of parameters, in the Rust shim (grape-vector-db_b200/rust/src/lib.rs) and in the C++ host mirror
(grape-vector-db_b200/host/gvdb_host.hpp).  No Rust toolchain exists in this image, so this is the check
that keeps the uncompiled crate honest; it reads /root/reference and is skipped where that tree is absent
(the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
RUST = os.path.join(ROOT, "grape-vector-db_b200", "rust", "src", "lib.rs")
HPP = os.path.join(ROOT, "grape-vector-db_b200", "host", "gvdb_host.hpp")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _block(text: str, header_re: str) -> str:
    """The brace-balanced body that follows the first match of header_re."""
    m = re.search(header_re, text)
    assert m, header_re
    i = text.index("{", m.end() - 1)
    depth, j = 0, i
    while True:
        c = text[j]
        depth += c == "{"
        depth -= c == "}"
        if depth == 0:
            return text[i + 1:j]
        j += 1


def _split_params(params: str):
    out, depth, cur = [], 0, ""
    for ch in params:
        if ch in "<([":
            depth += 1
        elif ch in ">)]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [p.strip() for p in out if p.strip()]


def _rust_fns(body: str, public_only: bool):
    """name -> (number of parameters besides self, takes &mut self, top-level only)."""
    fns = {}
    depth = 0
    pat = re.compile(r"(pub\s+)?fn\s+(\w+)\s*(<[^>]*>)?\s*\(")
    i = 0
    while i < len(body):
        c = body[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
        elif depth == 0:
            m = pat.match(body, i)
            if m and (m.group(1) or not public_only):
                j, d = m.end(), 1
                while d:
                    d += body[j] == "("
                    d -= body[j] == ")"
                    j += 1
                params = _split_params(body[m.end():j - 1])
                has_self = bool(params) and re.fullmatch(r"&?\s*(mut\s+)?self", params[0]) is not None
                fns[m.group(2)] = (len(params) - (1 if has_self else 0), has_self and "mut" in params[0])
                i = j
                continue
        i += 1
    return fns


def _cpp_methods(class_body: str):
    """name -> number of parameters, for the member functions declared at the top level of a class body."""
    body = re.sub(r"//[^\n]*", "", class_body)
    out = {}
    depth, i = 0, 0
    pat = re.compile(r"\b(\w+)\s*\(")
    skip = {"if", "for", "while", "return", "throw", "switch", "sizeof", "static_assert", "check"}
    while i < len(body):
        c = body[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
        elif depth == 0:
            m = pat.match(body, i)
            if m and m.group(1) not in skip and (i == 0 or not (body[i - 1].isalnum() or body[i - 1] == "_")):
                j, d = m.end(), 1
                while d:
                    d += body[j] == "("
                    d -= body[j] == ")"
                    j += 1
                out.setdefault(m.group(1), len(_split_params(body[m.end():j - 1])))
                i = j
                continue
        i += 1
    return out


def test_binary_quantizer_signatures_match_the_reference():
    ref = open(os.path.join(REF, "quantization.rs")).read()
    want = _rust_fns(_block(ref, r"impl\s+BinaryQuantizer\s*\{"), public_only=True)
    assert {"new", "quantize", "quantize_batch", "hamming_distance", "similarity", "multi_stage_search",
            "clear_cache", "get_cache_stats"} <= set(want), want
    rust = open(RUST).read()
    got = _rust_fns(_block(rust, r"impl\s+BinaryQuantizer\s*\{"), public_only=True)
    for name, (arity, mut_self) in want.items():
        assert name in got, f"rust shim lacks BinaryQuantizer::{name}"
        assert got[name] == (arity, mut_self), f"BinaryQuantizer::{name}: reference {(arity, mut_self)}, shim {got[name]}"
    hpp = open(HPP).read()
    cpp = _cpp_methods(_block(hpp, r"class\s+BinaryQuantizer\s*\{"))
    for name, (arity, _) in want.items():
        cname = "BinaryQuantizer" if name == "new" else name
        assert cname in cpp, f"C++ mirror lacks BinaryQuantizer::{name}"
        assert cpp[cname] == arity, f"C++ BinaryQuantizer::{name}: reference takes {arity} parameters, mirror {cpp[cname]}"


def test_binary_vector_from_bytes_call_matches_the_reference():
    ref = open(os.path.join(REF, "quantization.rs")).read()
    fns = _rust_fns(_block(ref, r"impl\s+BinaryVector\s*\{"), public_only=True)
    arity = fns["from_bytes"][0]
    rust = open(RUST).read()
    calls = re.findall(r"BinaryVector::from_bytes\(([^;]*?)\)\)?\s*[.;)]", rust)
    assert calls, "the shim builds its BinaryVector values with from_bytes"
    for c in calls:
        assert len(_split_params(c)) == arity, f"BinaryVector::from_bytes takes {arity} arguments in the reference: {c!r}"


def test_vector_index_trait_methods_match_the_reference():
    ref = open(os.path.join(REF, "index.rs")).read()
    want = _rust_fns(_block(ref, r"pub\s+trait\s+VectorIndex\s*:\s*Send\s*\+\s*Sync\s*\{"), public_only=False)
    assert set(want) == {"add_vector", "add_vectors", "search", "remove_vector", "len", "is_empty", "optimize",
                         "clear", "get_stats"}, want
    rust = open(RUST).read()
    got = _rust_fns(_block(rust, r"impl\s+VectorIndex\s+for\s+GpuVectorIndex\s*\{"), public_only=False)
    for name, sig in want.items():
        assert got.get(name) == sig, f"VectorIndex::{name}: reference {sig}, rust impl {got.get(name)}"
    hpp = open(HPP).read()
    cpp = _cpp_methods(_block(hpp, r"class\s+VectorIndex\s*\{"))
    for name, (arity, _) in want.items():
        assert cpp.get(name) == arity, f"C++ VectorIndex::{name}: reference takes {arity} parameters, mirror {cpp.get(name)}"


def test_ffi_declarations_cover_the_header():
    """Every GVDB_API function of include/gvdb.h that the Rust side declares has the header's parameter count."""
    hdr = open(os.path.join(ROOT, "include", "gvdb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {m.group(1): len(_split_params(m.group(2))) if m.group(2).strip() not in ("", "void") else 0
            for m in re.finditer(r"GVDB_API\s+[\w\s\*]+?\b(gvdb_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)}
    rust = open(RUST).read()
    ext = _block(rust, r'extern\s+"C"\s*\{')
    rfn = {m.group(1): len(_split_params(m.group(2)))
           for m in re.finditer(r"pub\s+fn\s+(gvdb_\w+)\s*\(([^;]*?)\)\s*(?:->[^;]+)?;", ext, flags=re.S)}
    assert len(rfn) >= 40
    for name, n in rfn.items():
        assert name in decl, f"{name} is declared in the Rust crate but not in include/gvdb.h"
        assert decl[name] == n, f"{name}: header has {decl[name]} parameters, Rust declaration {n}"
