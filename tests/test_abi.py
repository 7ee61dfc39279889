"""The C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every symbol
that include/tds_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tds_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tds_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER])


def test_library_exports_every_declared_symbol():
    from torchdrivesim_b200 import _build, _lib
    path = _build.build_library()
    lib = ctypes.CDLL(path)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in tds_b200.h but not exported"
    # the ctypes table binds exactly the declared functions
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().tds_version() >= 100


def test_struct_layouts_match_the_header():
    """sizeof of the ctypes mirrors == sizeof of the C structs (compiled with gcc)."""
    from torchdrivesim_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "tds_b200.h"
int main(void) { printf("%zu %zu %zu\n", sizeof(tds_kinematic_params_t), sizeof(tds_palette_t), sizeof(tds_map_info_t)); return 0; }
'''
    exe = os.path.join(ROOT, "tests", "_build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.KinematicParams), ctypes.sizeof(_lib.Palette), ctypes.sizeof(_lib.MapInfo)]


def test_sass_is_sm100a_only():
    from torchdrivesim_b200 import _build
    out = subprocess.run(["cuobjdump", "--list-elf", _build.build_library()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_invalid_arguments_are_reported_not_crashed():
    from torchdrivesim_b200 import _lib
    lib = _lib.load()
    p = _lib.KinematicParams(0.1, 5.0, 1.57, 1.57, 0, 20.0, 31.4, 5.0)
    # null pointers with n > 0 -> error code + message, no launch
    rc = lib.tds_kinematic_step_fwd(None, None, 2, None, None, 0, 10, ctypes.byref(p), None, None)
    assert rc == 1 and b"null" in lib.tds_last_error()
    rc = lib.tds_collision_allpairs_fwd(None, None, None, 1, 1, 1, 7, 1, None, None, None)
    assert rc == 1
    assert lib.tds_raster_workspace_bytes(-1, 0, 0, 0) == -1
    assert lib.tds_raster_workspace_bytes(2, 3, 1, 0) >= 2 * (11 * 24 + 11)
    with pytest.raises(_lib.TdsError):
        _lib.check(rc)
