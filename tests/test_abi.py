"""The C-ABI library: it loads without a GPU, exports every symbol include/tetra_b200.h declares,
and refuses to run (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

import tetra_testlib as T


def declared_functions():
    src = open(os.path.join(T.ROOT, "include", "tetra_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("tb200_create", "tb200_rx_stream_host", "tb200_rx_stream_dev", "tb200_expand_records",
                 "tb200_descramble_deinterleave", "tb200_decode_blocks", "tb200_find_train_seq"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()                           # incremental: rebuilds only when the sources are newer
    out = subprocess.check_output(["nm", "-D", "--defined-only", T.PRODUCT_SO], text=True)
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing
    lib = C.CDLL(T.PRODUCT_SO)          # loads without a GPU (static cudart)
    lib.tb200_version.restype = C.c_char_p
    assert b"sm_100a" in lib.tb200_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = C.CDLL(T.PRODUCT_SO)
    h = C.c_void_p()
    rc = lib.tb200_create(C.byref(h), 0)
    assert rc < 0 and not h.value       # fails loudly, nothing to fall back to


def test_record_expansion_order():
    """host-side expansion of slots into TMV-SAP records follows the reference's call order"""
    import numpy as np
    g = C.CDLL(T.PRODUCT_SO)
    g.tb200_expand_records.restype = C.c_size_t
    g.tb200_expand_records.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    slots = np.zeros(4, dtype=T.SLOT_DTYPE)
    slots["flags"] = [1 | 0x04 | 0x10, 2 | 0x04, 3 | 0x08, 0 | 0x20]
    slots["scrambling_code"] = 0x41802a07
    slots["time"] = 2 | (18 << 3) | (7 << 8)
    t1 = np.zeros((4, 288), np.uint8)
    t1[0, 60:74] = 1
    rec = np.zeros(16, dtype=T.RECORD_DTYPE)
    n = g.tb200_expand_records(slots.ctypes.data_as(C.c_void_p), t1.ctypes.data_as(C.c_void_p), 4,
                               rec.ctypes.data_as(C.c_void_p), 16)
    assert n == 8
    assert list(rec["lchan"][:8]) == [T.LC_BSCH, T.LC_AACH, T.LC_BNCH, T.LC_AACH, T.LC_SCH_F, T.LC_AACH, 0, 0]
    assert list(rec["blk_num"][:8]) == [1, 0, 2, 0, 0, 0, 1, 2]
    assert list(rec["type1_len"][:8]) == [60, 14, 124, 14, 268, 14, 124, 124]
    assert list(rec["crc_ok"][:8]) == [1, 1, 0, 1, 1, 1, 0, 1]
    assert rec["scrambling_code"][0] == 3 and rec["scrambling_code"][1] == 0x41802a07
    assert rec["type1"][1][:14].all() and (rec["tn"][0], rec["fn"][0], rec["mn"][0]) == (2, 18, 7)
