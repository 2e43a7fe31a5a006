"""C-ABI checks that need no GPU: the shared library loads, exports every symbol include/maf.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import mafb200 as maf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "maf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(maf_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = maf.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"libmembrane_b200.so does not export {s}"
    assert sorted(maf.pkg.capi.EXPORTS) == syms


def test_struct_layouts_match_header(tmp_path):
    """Compile include/maf.h with gcc and compare offsetof/sizeof of every field with the ctypes mirrors
    (a mismatch would silently corrupt every call)."""
    import subprocess
    capi = maf.pkg.capi
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "maf.h")}"',
             'int main(void) {']
    for cname, cls in (("maf_mesh_desc", capi.MeshDesc), ("maf_params", capi.ParamsC)):
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in (("maf_mesh_desc", capi.MeshDesc), ("maf_params", capi.ParamsC)):
        assert int(out[cname]) == C.sizeof(cls)
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, fname


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = maf.Params(motion=maf.LAG, scenario=maf.F_PULL, num1el=3, num2el=3, output=False)
    mesh = maf.Mesh(p, pull_speed=0.5)
    with pytest.raises(maf.MafError, match="no CUDA device"):
        maf.Assembler(mesh, p)
    with pytest.raises(maf.MafError):
        maf.calc_r_K(mesh, *maf.synthetic_state(mesh, p), 0.5, 0.5, p)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "membranealefem.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for line in src.splitlines():
                    if re.search(r"^\s*(from|import)\s+.*oracle|#include.*oracle|dlopen.*oracle|CDLL.*oracle", line):
                        raise AssertionError(f"{f}: the product path must not load test infrastructure: {line}")
