"""CPU: the C-ABI library loads, exports every symbol include/bsr.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from better_search_rag_rust_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bsr.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"BSR_API[^;]*?\b(bsr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    declared = _declared()
    assert len(declared) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported, sorted(set(declared) - exported)
    # nothing else leaks out of the library
    assert {s for s in exported if not s.startswith("bsr_")} == set()


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == _declared()
    L = _lib.lib()
    for name in _lib.SIGNATURES:
        assert getattr(L, name) is not None


def test_rust_and_cxx_bindings_cover_the_header():
    """The Rust crate (source only: no cargo in this image) and the C++ host mirror must bind the
    same entry points as the header, so that they cannot drift apart unnoticed."""
    declared = set(_declared())
    rust = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    rust_fns = set(re.findall(r"pub fn (bsr_[a-z0-9_]+)\s*\(", rust))
    assert rust_fns == declared, sorted(rust_fns ^ declared)
    hpp = open(os.path.join(ROOT, "better_search_rag_rust_b200", "csrc", "bsr_host.hpp")).read()
    used = set(re.findall(r"\b(bsr_[a-z0-9_]+)\s*\(", hpp))
    assert used <= declared | {"bsr_check"}, sorted(used - declared)
    # the search path proper is reachable from C++
    assert {"bsr_store_create", "bsr_store_append", "bsr_search", "bsr_search_parallel"} <= used


def test_header_compiles_as_c_and_cxx(tmp_path):
    for comp, ext, std in (("gcc", "c", "-std=c99"), ("g++", "cpp", "-std=c++17")):
        f = tmp_path / f"t.{ext}"
        f.write_text('#include "bsr.h"\nint main(void){ bsr_store *s = 0; (void)s; return BSR_OK; }\n')
        subprocess.check_call([comp, std, "-Wall", "-Werror", "-pedantic", "-I",
                               os.path.join(ROOT, "include"), "-c", str(f), "-o", str(tmp_path / "t.o")])


def test_constants_and_slots():
    L = _lib.lib()
    assert L.bsr_version() == 200
    assert [L.bsr_candidate_slots(k) for k in (1, 10, 32, 33, 50, 100, 1000, 1024)] == \
        [32, 32, 32, 64, 64, 128, 1024, 1024]
    assert L.bsr_candidate_slots(0) == 0 and L.bsr_candidate_slots(1025) == 0


def _no_gpu():
    n = C.c_int(-1)
    rc = _lib.lib().bsr_device_count(C.byref(n))
    return rc != 0 or n.value == 0


def test_argument_errors_do_not_need_a_gpu():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.bsr_store_create(0, 0, _lib.BSR_F32, 0, 0, C.byref(h)) == _lib.BSR_ERR_INVALID
    assert b"dim" in L.bsr_last_error()
    assert L.bsr_store_create(0, 768, 7, 0, 0, C.byref(h)) == _lib.BSR_ERR_INVALID
    assert L.bsr_store_create(0, 768, _lib.BSR_F32, 0, 0, None) == _lib.BSR_ERR_INVALID
    assert L.bsr_store_count(None, None) == _lib.BSR_ERR_INVALID
    assert L.bsr_search(None, None, 1, 10, None, None, None) == _lib.BSR_ERR_INVALID
    assert L.bsr_store_destroy(None) == _lib.BSR_OK


def test_fails_loudly_without_a_cuda_device():
    if not _no_gpu():
        pytest.skip("a CUDA device is visible")
    h = C.c_void_p()
    rc = _lib.lib().bsr_store_create(0, 768, _lib.BSR_F32, 0, 0, C.byref(h))
    assert rc == _lib.BSR_ERR_CUDA and not h.value
    import better_search_rag_rust_b200 as bsr
    with pytest.raises(bsr.BsrError) as e:
        bsr.GpuVectorstore(None, True, dim=768)
    assert e.value.status == _lib.BSR_ERR_CUDA
    with pytest.raises(bsr.BsrError):
        s = bsr.GpuVectorstore(None, True)
        s.append_many([[1.0, 2.0, 3.0]])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "better_search_rag_rust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "bsr_oracle_" not in text or f == "bsr_kernels.cuh", f
