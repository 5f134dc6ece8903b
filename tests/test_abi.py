"""The C-ABI shared library: builds for sm_100a, loads, exports every symbol include/*.h declares, and refuses to
run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import iak3d_b200
from iak3d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for f in os.listdir(os.path.join(ROOT, "include")):
        if f.endswith(".h"):
            src = open(os.path.join(ROOT, "include", f)).read()
            names += re.findall(r"^IAK3D_API\s+[A-Za-z_0-9 \*]+?\b(iak3d_[A-Za-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_library_exists_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `make` / __graft_entry__.build() first"
    _lib.load()


def test_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert sorted(_lib.EXPORTS) == declared, "the ctypes prototypes and the header disagree"


def test_pars_struct_layout(tmp_path):
    """struct iak3d_pars as gcc lays it out from include/iak3d.h == the ctypes mirror, field by field."""
    import shutil
    import subprocess
    fields = [f[0] for f in _lib.Pars._fields_]
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    src = tmp_path / "layout.c"
    body = "".join(f'  printf("{f} %zu\\n", offsetof(iak3d_pars, {f}));\n' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "iak3d.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(iak3d_pars));\n' + body + '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == C.sizeof(_lib.Pars)
    for f in fields:
        assert int(out[f]) == getattr(_lib.Pars, f).offset, f
    assert _lib.MAX_SPL == 12


def test_r_shim_compiles_against_the_header():
    """r/r_shim.c (the .Call binding a maintainer builds where R exists) must agree with include/iak3d.h: syntax / type
    check with gcc against minimal stand-ins of the R headers (tests/r_stub; R itself is not in this image)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "r_stub"), "-I",
                        os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "r_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "r", "r_shim.c")).read()
    glue = open(os.path.join(ROOT, "r", "iak3d_cuda.R")).read()
    for name in re.findall(r"\.Call\('(iak3dR_[A-Za-z0-9_]+)'", glue):
        assert '{"%s"' % name in src, f"{name} is called from the R glue but not registered in r_shim.c"


def test_sass_is_sm100a_with_tma_and_dmma():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "DMMA" in out, "the factorisation must use the FP64 tensor pipe"
    assert "UBLKCP" in out, "operands must be staged by TMA bulk copies"


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(iak3d_b200.Iak3dError):
        iak3d_b200.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "iak3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
