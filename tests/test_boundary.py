"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares; no compute call works without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import sse_ffi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(REPO, "include", "sse_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sse_[a-z0-9_]+)\s*\(", src)))


def test_library_built_and_exports_all_header_symbols():
    assert os.path.exists(sse_ffi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(sse_ffi.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(sse_ffi.EXPORTED_SYMBOLS) == names, "sse_ffi signatures out of sync with the header"
    lib.sse_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sse_version()


def test_config_struct_size_checked():
    lib = sse_ffi.load_library()
    cfg = sse_ffi.SseConfig()
    cfg.struct_size = 4
    h = ctypes.c_void_p()
    assert lib.sse_create(ctypes.byref(cfg), ctypes.byref(h)) < 0
    assert b"struct_size" in lib.sse_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sse_ffi.SseError) as e:
        sse_ffi.Handle("dual-encoder", 100, 8, 8, 8, 8, 10)
    assert "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(sse_ffi.SseError):
        sse_ffi.load_library(str(tmp_path / "nope.so"))


def test_header_is_valid_c_and_links_from_a_c_program(tmp_path):
    """The boundary is a C ABI: include/sse_b200.h must compile as C99 and a plain C program must link against the
    library and call (host-only) entry points."""
    import subprocess
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(repo, "include", "sse_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    src = tmp_path / "use.c"
    src.write_text('#include "sse_b200.h"\n#include <stdio.h>\n#include <string.h>\nint main(void) {\n'
                   '  float v[3] = {0.1f, 1e-5f, 1234567.0f}; char out[64]; int64_t ends[3];\n'
                   '  if (sse_tsv_format_f32(v, 3, out, sizeof out, ends) != 0) return 1;\n'
                   '  printf("%.*s %08x\\n", (int)ends[2], out, sse_crc32c("123456789", 9, 0));\n'
                   '  return strlen(sse_version()) == 0;\n}\n')
    libdir = os.path.join(repo, "sequence-semantic-embedding_b200")
    exe = str(tmp_path / "use_c")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(repo, "include"), str(src), "-o", exe, "-L", libdir, "-l:libsse_b200.so",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["0.11e-051.234567e+06", "e3069283"]
