"""clover_b200 - B200-native (sm_100a) implementation of the Clover quantized linear-algebra hot path.

Layers:
  csrc/            hand-written CUDA kernels + the C ABI  ->  libclover_b200.so   (the product)
  _lib.py          ctypes binding of include/clover_b200.h (fails loudly when the library is missing)
  containers.py    device-resident mirror of the reference containers (CloverVector4, CloverMatrix4, ...)
  sharded.py       row-sharded multi-GPU mvm over torch.distributed / NCCL
  apps.py          the reference's application loops (Q_IHT, Q_GD) composed from the containers
"""
from ._lib import (ABI_SYMBOLS, DOT_AUTO, DOT_EXACT, DOT_FAST, THRESHOLD_AUTO, THRESHOLD_EXACT, THRESHOLD_FAST,  # noqa: F401
                   CloverError, build, call, lib)


def __getattr__(name):
    # containers import torch; keep `import clover_b200` light for the ABI/symbol tests
    if name in ("CloverVector32", "CloverVector4", "CloverVector8", "CloverMatrix32", "CloverMatrix4",
                "CloverMatrix8", "CloverSizeError", "size_pad"):
        from . import containers
        return getattr(containers, name)
    raise AttributeError(name)
