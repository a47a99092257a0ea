"""osmo-tetra_b200 - B200-native TETRA lower-MAC receive chain (type-5 bits -> type-1 bits).

The product is the CUDA library `libtetra_b200.so` built from csrc/ (C ABI in
include/tetra_b200.h); `host/` holds the C side that plugs it into the reference
(tetra_shim.c, tetra_rx_b200.c).  `binding.py` is the ctypes face of the C ABI (class B200,
the record layouts, the multi-GPU sharding driver); this module finds the in-tree library and
refuses to continue without it - there is no CPU implementation behind this package.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtetra_b200.so")


def load_library():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    return ctypes.CDLL(LIB_PATH)
