"""CPU: libtrx.so loads and exports every symbol include/trx.h declares; failures are reported
through return codes and trx_last_error, never by aborting.  No compute happens here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "trx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trx_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for need in ("trx_create", "trx_add", "trx_search", "trx_set_groups", "trx_reset", "trx_destroy",
                 "trx_last_error", "trx_stats", "trx_merge_topk"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from textreact_b200 import _lib
    L = _lib.lib()
    for s in declared_symbols():
        assert hasattr(L, s), f"libtrx.so does not export {s}"
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert b"sm_100a" in L.trx_version()


def test_stats_struct_matches_header_layout(tmp_path):
    """ctypes mirror of trx_stats_t / trx_search_params_t == what a C compiler makes of include/trx.h."""
    import subprocess
    from textreact_b200 import _lib
    fields = [n for n, _ in _lib.TrxStats._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "trx.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(trx_stats_t));\n'
                   + "".join(f'  printf("%zu\\n", offsetof(trx_stats_t, {n}));\n' for n in fields)
                   + '  printf("%zu\\n", sizeof(trx_search_params_t));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert got[0] == ctypes.sizeof(_lib.TrxStats)
    assert got[1:-1] == [getattr(_lib.TrxStats, n).offset for n in fields]
    assert got[-1] == ctypes.sizeof(_lib.TrxSearchParams)


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import textreact_b200 as trx
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        trx.IndexFlatIP(8)
    from textreact_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.trx_create(8, 0, 0, ctypes.byref(h)) == _lib.TRX_ENODEV
    assert not h
    assert L.trx_create(-1, 0, 0, ctypes.byref(h)) == _lib.TRX_EINVAL
    assert b"dimension" in L.trx_last_error()
    assert L.trx_search(None, None, 1, 1, None, None, None, None) == _lib.TRX_EINVAL


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "textreact_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} mentions the oracle"


def _compile_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "trx_example")
    lib = os.path.join(ROOT, "textreact_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", lib, "-ltrx",
                           f"-Wl,-rpath,{lib}", "-o", exe])
    return exe


def test_header_is_plain_c_and_the_example_links(tmp_path):
    """include/trx.h compiles as C99 with -Wall -Wextra -Werror and examples/c_abi_example.c links against libtrx.so."""
    import subprocess
    import torch
    exe = _compile_example(tmp_path)
    if not torch.cuda.is_available():
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 2 and "no CPU fallback" in p.stderr          # trx_create -> TRX_ENODEV, reported not aborted


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    import subprocess
    exe = _compile_example(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "query 0: (0, 0.0000)" in p.stdout and "ntotal=20000" in p.stdout
