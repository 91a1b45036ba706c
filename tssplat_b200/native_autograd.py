"""Optional C++ autograd bridge (``csrc/torch_binding.cpp``): the reference's pybind11-module route of
INTEGRATION.md section 2, built with ``torch.utils.cpp_extension`` in-tree (``tssplat_b200/_torch_build``).

It removes the Python interpreter from between autograd and the two launches of an iteration (forward: the fused
kernel; backward: ``tsb_scale``); semantics are those of ``tet_spheres_ext.forward/backward``.  When the module has not
been built (or cannot be loaded) ``available()`` is False and ``tssplat_b200.energies`` keeps using the Python
``torch.autograd.Function`` -- both routes end in the same CUDA launches, neither is a CPU path.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_torch_build")
NAME = "_tsb_torch"
_mod = None
_tried = False


def build(verbose: bool = False) -> Optional[str]:
    """Compile the binding (g++ via ninja; no CUDA sources).  Returns the .so path or None on failure."""
    try:
        import torch  # noqa: F401
        from torch.utils import cpp_extension as ce
        os.makedirs(BUILD_DIR, exist_ok=True)
        env_cc = {k: os.environ.pop(k) for k in ("CC", "CXX") if k in os.environ}      # this image's $CC wrapper is not a C++ driver
        try:
            ce.load(name=NAME, sources=[os.path.join(HERE, "csrc", "torch_binding.cpp")], build_directory=BUILD_DIR,
                    extra_cflags=["-O2", "-std=c++17"], with_cuda=True, is_python_module=False, verbose=verbose)
        finally:
            os.environ.update(env_cc)
        so = os.path.join(BUILD_DIR, NAME + ".so")
        return so if os.path.exists(so) else None
    except Exception as ex:  # the Python route stays available
        if verbose:
            print(f"native autograd bridge not built: {type(ex).__name__}: {ex}", file=sys.stderr)
        return None


def _load():
    global _mod, _tried
    if _tried:
        return _mod
    _tried = True
    so = os.path.join(BUILD_DIR, NAME + ".so")
    if os.environ.get("TSSPLAT_B200_NO_NATIVE_AUTOGRAD") or not os.path.exists(so):
        return None
    try:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        spec = importlib.util.spec_from_file_location(NAME, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        from . import _capi
        addr = lambda f: C.cast(f, C.c_void_p).value      # noqa: E731
        mod.bind(addr(_capi.lib.tsb_energy_grad), addr(_capi.lib.tsb_scale), addr(_capi.lib.tsb_last_error))
        _mod = mod
    except Exception:
        _mod = None
    return _mod


def available() -> bool:
    return _load() is not None


def module():
    return _load()
