"""CPU-only: the C-ABI shared library loads without a GPU and exports every function that
include/graphnics_b200.h declares; the ctypes struct mirrors have the C layout."""
import ctypes as C
import os
import re
import subprocess

from _util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "graphnics_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from graphnics_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert n in _lib.PROTOTYPES, f"{n} declared in the header but not bound in _lib.PROTOTYPES"
        assert getattr(lib, n) is not None
    for n in _lib.PROTOTYPES:
        assert n in names, f"{n} bound in _lib.PROTOTYPES but not declared in the header"
    assert lib.gnx_version() >= 1
    assert lib.gnx_device_ok() in (0, 1)


def test_struct_layouts_match_c(tmp_path):
    from graphnics_b200 import _lib
    from graphnics_b200.coefficients import gnx_program_t
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "graphnics_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(gnx_network_t),sizeof(gnx_mixed_coeffs_t),sizeof(gnx_schur_plan_t),sizeof(gnx_program_t),'
                    'sizeof(gnx_exchange_t),sizeof(gnx_primal_coeffs_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_lib.gnx_network_t), C.sizeof(_lib.gnx_mixed_coeffs_t),
                     C.sizeof(_lib.gnx_schur_plan_t), C.sizeof(gnx_program_t), C.sizeof(_lib.gnx_exchange_t),
                     C.sizeof(_lib.gnx_primal_coeffs_t)]


def test_compute_entry_points_fail_loudly_without_gpu():
    import pytest
    import torch
    from graphnics_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GnxError):
        _lib.require_device()
