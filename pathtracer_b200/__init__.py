"""pathtracer_b200 — B200-native radiance loop for nbonneel/pathtracer behind a C-ABI.

The product is `csrc/libptb200.so` (hand-written CUDA for sm_100a + a C++ host side that builds the
wide BVH).  This package is the thin host mirror of the reference's `Raytracer` interface over that
library.  There is no CPU path: if the library is not built, `load()` raises.
"""
import ctypes
import os

from . import _abi
from ._abi import Lib, PtbError  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTB_LIB_PATH: A/B measurements of another build of the same library (experiments only)
LIB_PATH = os.environ.get("PTB_LIB_PATH") or os.path.join(_HERE, "csrc", "libptb200.so")
_lib = None
_io = None


def load():
    """Bind libptb200.so (built in-tree by __graft_entry__.build() / `make -C pathtracer_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PtbError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback.")
        _lib = Lib(ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL), "ptb_")
    return _lib


def sceneio():
    """Bind the scene-file readers of include/ptb_sceneio.h (same library; host code, usable without a GPU)."""
    global _io
    if _io is None:
        _io = _abi.SceneIO(load().cdll, "ptb_")
    return _io


def Raytracer(device=0):
    """A reference-shaped `Raytracer` bound to the CUDA library."""
    from .api import Raytracer as _R
    return _R(load(), device)
