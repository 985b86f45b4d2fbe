"""The C-ABI shared library loads and exports every symbol include/cnnb200.h declares; the product
package never touches the oracle; without a GPU the product fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cnnb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cnnb200_[a-z0-9_]+)\s*\(", text)) - {"cnnb200_alloc_fn", "cnnb200_free_fn"})


def test_library_builds_and_exports_every_declared_symbol():
    from commonnnclustering_b200 import build

    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/cnnb200.h but not exported"


def test_header_is_plain_c_and_links(tmp_path):
    """A foreign-function binding (cgo, Cython `cdef extern`, ctypes generators) reads the header as C, not C++: a
    strict C99 translation unit that takes the address of EVERY declared entry point compiles without a warning, links
    against the library and runs -- without a GPU `cnnb200_ctx_create` answers with an error code and a message, it
    does not crash and nothing falls back."""
    from commonnnclustering_b200 import build

    lib_path = build.build()
    syms = declared_symbols()
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "cnnb200.h"\n'
        "typedef void (*any_fn)(void);\n"
        "int main(void) {\n"
        "    any_fn table[] = {" + ", ".join(f"(any_fn){name}" for name in syms) + "};\n"
        "    cnnb200_ctx *ctx = NULL;\n"
        "    int rc = cnnb200_ctx_create(0, &ctx);\n"
        '    printf("%d %d %d %s\\n", (int)(sizeof table / sizeof table[0]), cnnb200_version(), rc, rc ? cnnb200_last_error() : "ok");\n'
        "    if (rc == 0) cnnb200_ctx_destroy(ctx);\n"
        "    return 0;\n}\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                    "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), lib_path,
                    f"-Wl,-rpath,{os.path.dirname(lib_path)}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split(None, 3)
    assert int(out[0]) == len(syms) and int(out[1]) == 100
    assert (int(out[2]) == 0 and out[3].strip() == "ok") or (int(out[2]) < 0 and len(out[3].strip()) > 0)


def test_python_binding_covers_the_header():
    from commonnnclustering_b200 import _lib

    bound = set(_lib.SIGNATURES) | {"cnnb200_version", "cnnb200_last_error"}
    assert set(declared_symbols()) == bound


def test_version_and_error_string():
    from commonnnclustering_b200 import _lib

    lib = _lib.load()
    assert lib.cnnb200_version() == 100
    assert isinstance(lib.cnnb200_last_error(), bytes)


def test_sass_is_sm100a():
    from commonnnclustering_b200 import build

    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sass_shows_the_tcgen05_tma_path():
    """The tiled distance kernel really is tcgen05 / TMEM / TMA code (and the committed summary matches the library):
    UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor."""
    import importlib.util

    from commonnnclustering_b200 import build

    build.build()
    spec = importlib.util.spec_from_file_location("sass_summary", os.path.join(ROOT, "scripts", "sass_summary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    total, per_kernel, arch = mod.counts()
    assert arch == ["sm_100a"]
    assert total["UTCHMMA"] >= 8 and total["LDTM"] >= 4 and total["UTMALDG"] >= 8 and total["SYNCS"] >= 8
    # one instantiation per number of 64-wide K atoms (1 .. 4): 8 MMAs (2 row tiles x 4 K steps) and 3 TMA loads
    # (2 A tiles + the B stage) per atom, four tcgen05.ld per accumulator tile
    tensor = sorted((v for k, v in per_kernel.items() if "tensor_tiles_kernel" in k), key=lambda v: v.get("UTCHMMA", 0))
    assert len(tensor) == 4
    for atoms, v in enumerate(tensor, start=1):
        assert v.get("UTCHMMA", 0) == 8 * atoms and v.get("LDTM", 0) >= 4 and v.get("UTMALDG", 0) == 3 * atoms
    committed = open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read()
    for m in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert re.search(rf"^{m}\s+{total[m]}\b", committed, flags=re.M), m


def test_product_never_imports_the_oracle():
    # the package, the public header and the measurement scripts: only tests/ (incl. tests/checks/), smoke() and
    # bench.py's cpu_baseline / reference legs may touch oracle/
    for top in ("commonnnclustering_b200", "include", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                    assert "cnn_oracle" not in text, f
    code = "import sys; import commonnnclustering_b200.cluster, commonnnclustering_b200._lib; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from commonnnclustering_b200 import _lib

    with pytest.raises(Exception) as e:
        _lib.Context(0)
    assert "cnnb200 error" in str(e.value)


def test_missing_library_raises_import_error(tmp_path, monkeypatch):
    from commonnnclustering_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()
