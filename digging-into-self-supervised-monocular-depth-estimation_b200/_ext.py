"""Loader of the native extension.  There is no fallback: if the CUDA library or the
C++ extension is missing this raises, and every op of the package raises with it."""
from __future__ import annotations

import importlib.util
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_EXT_PATH = os.path.join(_PKG, "_md2_torch.so")
_LIB_PATH = os.path.join(_PKG, "libmd2loss.so")
_ext = None


class NativeExtensionMissing(RuntimeError):
    pass


def ext():
    global _ext
    if _ext is None:
        for p in (_LIB_PATH, _EXT_PATH):
            if not os.path.exists(p):
                raise NativeExtensionMissing(
                    f"{p} is missing. This package has no CPU or PyTorch fallback for the fused loss; "
                    "build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        spec = importlib.util.spec_from_file_location("_md2_torch", _EXT_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ext = mod
    return _ext
