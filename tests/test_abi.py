"""The C-ABI library: builds, loads, exports every symbol include/ptb200.h declares, fails loudly without a GPU,
and contains no host-side instantiation of the device code (no CPU path)."""
import ctypes
import os
import re
import subprocess

import pytest

import pathtracer_b200
from pathtracer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted("ptb_" + s for s in _abi.SYMBOLS + _abi.MULTI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = pathtracer_b200.load()          # raises if libptb200.so is not built: there is no fallback
    out = subprocess.check_output(["nm", "-D", "--defined-only", pathtracer_b200.LIB_PATH], text=True)
    exported = set(re.findall(r" T (ptb_\w+)", out))
    assert set(header_functions()) <= exported
    assert lib.version().decode().startswith("ptb200")


def test_library_has_no_host_copy_of_the_device_code():
    out = subprocess.check_output(["nm", "-C", "--defined-only", pathtracer_b200.LIB_PATH], text=True)
    host_syms = [l for l in out.splitlines() if re.search(r" [TtWw] ptb::(traverse|shade_one|extend_one|shadow_one|splat_pixel|raygen_one)", l)]
    assert host_syms == [], host_syms


def test_sm100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "--list-elf", pathtracer_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_sizes_match_the_header():
    code = r'''
    #include <stdio.h>
    #include "ptb200.h"
    int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ptb_tex), sizeof(ptb_material), sizeof(ptb_xform), sizeof(ptb_mesh),
                      sizeof(ptb_camera), sizeof(ptb_params), sizeof(ptb_stats), sizeof(ptb_scene_info)); return 0;}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")], text=True).split()]
    expect = [ctypes.sizeof(t) for t in (_abi.Tex, _abi.Material, _abi.Xform, _abi.Mesh, _abi.Camera, _abi.Params, _abi.Stats, _abi.SceneInfo)]
    assert sizes == expect


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = pathtracer_b200.load()
    ctx = ctypes.c_void_p()
    rc = lib.create(0, ctypes.byref(ctx))
    assert rc == -3 and not ctx.value
    assert b"no CPU path" in lib.last_error(None)
