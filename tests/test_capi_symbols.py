"""The C-ABI library builds, loads and exports every symbol that include/*.h declares.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, macro):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(macro + r"\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    names = declared("trueno_rag_b200.h", "TRR_API") + declared("trueno_rag_host.h", "TRRH_API")
    assert len(names) > 50
    for n in names:
        assert hasattr(built_lib, n), f"{n} is declared in include/ but not exported"


def test_prototype_tables_cover_the_headers(built_lib):
    from trueno_rag_b200 import _lib
    assert set(declared("trueno_rag_b200.h", "TRR_API")) == set(_lib.TRR_PROTOS)
    assert set(declared("trueno_rag_host.h", "TRRH_API")) == set(_lib.TRRH_PROTOS)


def test_version_and_error_string(built_lib):
    assert built_lib.trr_version() >= 100
    assert isinstance(built_lib.trr_last_error(), bytes)
    assert built_lib.trr_exchange_bytes(1024, 50) == (4 * 1024 * 50 + 2 * 1024) * 4


def test_no_cpu_fallback_without_a_device(built_lib):
    if built_lib.trr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert built_lib.trr_ctx_create(0, C.byref(ctx)) == 7  # TRR_ERR_NO_DEVICE
    assert b"no CPU fallback" in built_lib.trr_last_error()


def test_product_never_references_the_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "trueno_rag_b200")):
        if os.path.basename(root) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(root, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|libtrr_oracle|orc_\w+\s*\(", txt):
                    bad.append(f)
    assert not bad, bad


def test_rust_sys_crate_is_generated_from_the_header():
    """integration/rust/trueno-rag-b200-sys/src/lib.rs is the mechanical translation of include/trueno_rag_b200.h: one
    `pub fn` per TRR_API symbol, and regenerating it changes nothing"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rs = open(os.path.join(ROOT, "integration", "rust", "trueno-rag-b200-sys", "src", "lib.rs")).read()
    assert sorted(re.findall(r"pub fn (\w+)\(", rs)) == declared("trueno_rag_b200.h", "TRR_API")


def _top_level_args(src, i):
    """src[i] is the '(' of a call: returns the number of top-level arguments and the index after the closing ')'"""
    depth, n, seen = 0, 0, False
    j = i
    while j < len(src):
        c = src[j]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                return (n + 1 if seen else 0), j + 1
        elif c == "," and depth == 1:
            n += 1
            seen_after = src[j + 1:].lstrip()
            if seen_after.startswith(")"):  # trailing comma
                n -= 1
        elif not c.isspace() and depth >= 1:
            seen = True
        j += 1
    raise AssertionError("unbalanced call")


def test_rust_wrapper_calls_match_the_generated_signatures():
    """the safe wrapper crate cannot be compiled here (no rustc): at least every `sys::trr_*` call in it must name a
    function of the generated -sys crate and pass as many arguments as the C declaration has"""
    rs = open(os.path.join(ROOT, "integration", "rust", "trueno-rag-b200-sys", "src", "lib.rs")).read()
    arity = {}
    for m in re.finditer(r"pub fn (\w+)\(", rs):
        n, _ = _top_level_args(rs, m.end() - 1)
        arity[m.group(1)] = n
    shim = open(os.path.join(ROOT, "integration", "rust", "trueno-rag-b200", "src", "lib.rs")).read()
    shim = re.sub(r"//[^\n]*", "", shim)
    calls = list(re.finditer(r"sys::(trr_\w+)\(", shim))
    assert len(calls) >= 12
    for m in calls:
        name = m.group(1)
        assert name in arity, f"{name} is not in the -sys crate"
        n, _ = _top_level_args(shim, m.end() - 1)
        assert n == arity[name], f"{name}: the wrapper passes {n} arguments, the C ABI takes {arity[name]}"
    for const in set(re.findall(r"sys::(TRR_\w+)", shim)):
        assert re.search(rf"pub const {const}: c_int", rs), const
