"""Golden vector for the symbol slicer, written by the REFERENCE's own float_to_bits program
(oracle/_ref/float_to_bits, compiled unmodified from /root/reference/src/float_to_bits.c).
Run in the build container only:  python tests/golden/make_golden_slicer.py
  slicer.npz   4096 float32 symbols (random, all decision edges, +-0, inf, NaN, denormals) and the
               8192 bits the reference program wrote for them."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import tetra_testlib as T  # noqa: E402


def main():
    rng = np.random.default_rng(0x7E7A00F1)
    sym = (rng.standard_normal(4096) * 2.5).astype(np.float32)
    special = np.array([2.0, -2.0, 0.0, -0.0, np.nextafter(np.float32(2), np.float32(3)), np.nextafter(np.float32(2), np.float32(0)),
                        np.nextafter(np.float32(-2), np.float32(-3)), np.nextafter(np.float32(-2), np.float32(0)),
                        np.float32(1e-45), np.float32(-1e-45), np.inf, -np.inf, np.nan, 5.0, -5.0, 4.9999995, 1e30, -1e30],
                       dtype=np.float32)
    sym[:special.size] = special
    sym[100:100 + special.size] = special[::-1]
    with tempfile.TemporaryDirectory() as d:
        bits = T.ref_float_to_bits(sym, d)
    assert bits.size == 2 * sym.size
    np.savez_compressed(os.path.join(HERE, "slicer.npz"), sym=sym, bits=np.packbits(bits))
    print("slicer.npz:", sym.size, "symbols")


if __name__ == "__main__":
    main()
