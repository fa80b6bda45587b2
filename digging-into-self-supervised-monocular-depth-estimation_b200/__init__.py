"""md2_b200 - B200-native (sm_100a) view-synthesis training loss of monodepth2, a drop-in
for the loss path of russellgeum/Digging-into-Self-Supervised-Monocular-Depth-Estimation.

Public surface
  functional.view_synthesis_loss   fused warp + photometric + auto-mask + smoothness, fwd+bwd
  functional.param2matrix          pose parametrisation kernel (warp.py:126-153)
  compute.compute                  L2 drop-in for model_tool/processor.py:139-218
  cabi                             ctypes mirror of include/md2_loss.h
  build                            in-tree nvcc / g++ build of the native code
  synthetic                        KITTI-shaped synthetic batches

The CUDA extension is mandatory: nothing here falls back to PyTorch ops or the CPU.
"""
from . import cabi, synthetic  # noqa: F401

__all__ = ["functional", "compute", "cabi", "build", "synthetic"]


def __getattr__(name):
    if name in ("functional", "compute", "build", "modules", "metrics", "pipeline", "selfcheck", "trainer"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
