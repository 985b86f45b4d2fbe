#!/usr/bin/env python
"""Build the UNMODIFIED reference (cnnclustering 0.5.0) into ``oracle/_ref/``.

TEST INFRASTRUCTURE — not product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import what this script produces.

The reference's hot path lives in five Cython modules (``setup.py:31-43`` of the
reference).  This recipe does *not* run the reference's own build system and copies
no reference source into the repository:

* every ``src/cnnclustering/*.pyx`` is translated where it lies (read-only) by the
  ``cython`` compiler into ``oracle/_ref/build/<mod>.cpp`` and compiled with ``g++``
  into ``oracle/_ref/cnnclustering/<mod>.<abi>.so``;
* the three pure-Python modules the extension modules import at run time
  (``__init__``, ``recipes``, ``report``) are compiled to code objects and stored as
  marshalled ``.code`` files next to the ``.so`` files (``.pyc`` files do not travel to the GPU
  box); ``oracle/ref.py`` loads them.

``oracle/_ref/`` is git-ignored (build output only) but travels to the GPU box.

Compiler directives: those of the reference's ``setup.py:45-54`` with two changes that do
not touch any source line and do not change any result (same labels, checked):
``linetrace=False`` and ``legacy_implicit_noexcept=True``.  Under the Cython 0.29 the
reference was written for (``pyproject.toml:5``) neither a profiling hook nor an
exception check is emitted for ``cdef ... nogil`` functions; Cython 3 emits both unless told
otherwise, which makes the as-shipped build ~8x slower than the published timings.
"""
import os
import marshal
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CNN_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "cnnclustering")
BUILD = os.path.join(OUT, "build")

MODULES = ["_primitive_types", "_types", "_fit", "_bundle", "cluster"]
PURE = ["__init__", "recipes", "report"]

DIRECTIVES = {
    "language_level": "3",
    "embedsignature": "True",
    "boundscheck": "False",
    "wraparound": "False",
    "cdivision": "True",
    "nonecheck": "False",
    "linetrace": "False",
    "annotation_typing": "False",
    "legacy_implicit_noexcept": "True",
}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "cnnclustering"))


def built() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(PKG, m + suffix)) for m in MODULES) and all(
        os.path.exists(os.path.join(PKG, p + ".code")) for p in PURE
    )


def _build_one(mod: str) -> None:
    import numpy as np

    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pyx = os.path.join(REF_SRC, "cnnclustering", mod + ".pyx")
    cpp = os.path.join(BUILD, mod + ".cpp")
    so = os.path.join(PKG, mod + suffix)
    cmd = [sys.executable, "-m", "cython", "--cplus", "-I", REF_SRC, "-o", cpp]
    for k, v in DIRECTIVES.items():
        cmd += ["-X", f"{k}={v}"]
    cmd.append(pyx)
    subprocess.run(cmd, check=True, cwd=REF_SRC, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cxx = os.environ.get("CXX", "g++")
    subprocess.run(
        [
            cxx, "-O2", "-std=c++14", "-fPIC", "-shared", "-w",
            "-fno-var-tracking-assignments",
            "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
            "-I", sysconfig.get_paths()["include"], "-I", np.get_include(),
            cpp, "-o", so,
        ],
        check=True,
    )


def build(force: bool = False) -> bool:
    """Build into oracle/_ref; returns True when the reference is importable afterwards."""
    if not available():
        return built()
    if built() and not force:
        return True
    os.makedirs(PKG, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(MODULES)) as ex:
        list(ex.map(_build_one, MODULES))
    for p in PURE:
        src = os.path.join(REF_SRC, "cnnclustering", p + ".py")
        with open(src) as fp:
            code = compile(fp.read(), f"<reference>/cnnclustering/{p}.py", "exec")
        with open(os.path.join(PKG, p + ".code"), "wb") as fp:
            marshal.dump(code, fp)
    shutil.rmtree(BUILD, ignore_errors=True)  # generated C++ is not needed once the modules are linked
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
