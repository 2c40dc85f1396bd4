"""CPU-only checks of the C-ABI shared library: it loads, exports every symbol include/bsx.h declares,
and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bsx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from blobstreamx_b200 import lib

    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    so = ctypes.CDLL(lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(so, s)]
    assert not missing, missing
    assert so.bsx_version() == 100


def test_no_cpu_fallback():
    import torch
    from blobstreamx_b200 import lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.BsxError):
        lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "blobstreamx_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").lower() or "import oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "bsx_oracle" not in txt, f


def test_rust_ffi_crate_matches_header():
    """bindings/rust/bsx-sys/src/lib.rs is generated from include/bsx.h: the committed file is up to date and declares
    every entry point the header does."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "scripts", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, n = gen.main()
    assert open(gen.PATH).read() == text, "run scripts/gen_rust_ffi.py"
    syms = _declared_symbols()
    assert n == len(syms)
    for s_ in syms:
        assert f"pub fn {s_}(" in text, s_


def test_struct_layouts_match_the_bindings(tmp_path):
    """sizeof / offsetof of every struct of include/bsx.h as gcc lays them out == the numpy dtypes and ctypes structures
    the Python binding fills (a drifted field would silently shift every input that follows it)."""
    import subprocess
    import numpy as np
    from blobstreamx_b200 import inputs as I, lib

    structs = {"bsx_header_in": lib.HEADER_IN, "bsx_skip_in": lib.SKIP_IN, "bsx_step_in": lib.STEP_IN,
               "bsx_header_fields": I.HEADER_FIELDS_DTYPE, "bsx_commit_in": I.COMMIT_DTYPE, "bsx_commit_sig_in": I.COMMIT_SIG_DTYPE}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bsx.h"', "int main(void) {"]
    for name, dt in structs.items():
        lines.append(f'printf("{name} size %zu\\n", sizeof({name}));')
        for f in dt.names:
            lines.append(f'printf("{name} {f} %zu\\n", offsetof({name}, {f}));')
    for name, cls in (("bsx_skip_batch", lib.SkipBatch), ("bsx_range_batch", lib.RangeBatch)):
        lines.append(f'printf("{name} size %zu\\n", sizeof({name}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{name} {f} %zu\\n", offsetof({name}, {f}));')
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for name, dt in structs.items():
        assert got[(name, "size")] == np.dtype(dt).itemsize, name
        for f in dt.names:
            assert got[(name, f)] == dt.fields[f][1], (name, f)
    for name, cls in (("bsx_skip_batch", lib.SkipBatch), ("bsx_range_batch", lib.RangeBatch)):
        assert got[(name, "size")] == ctypes.sizeof(cls), name
        for f, _ in cls._fields_:
            assert got[(name, f)] == getattr(cls, f).offset, (name, f)
