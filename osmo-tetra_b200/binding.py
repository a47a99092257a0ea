"""ctypes binding of include/tetra_b200.h (libtetra_b200.so): the Python face of the C ABI.

Plumbing only - device buffers come from torch tensors (`tensor.data_ptr()`), results are numpy
structured arrays with the layouts of `struct tb200_slot` / `struct tb200_record`.  The tests load the
SIMT-emulation build of the same sources through the same class (`B200(lib_path=...)`).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.path.join(HERE, "libtetra_b200.so")

TS_NORM_1, TS_NORM_2, TS_NORM_3, TS_SYNC, TS_EXT = 0, 1, 2, 3, 4          # enum tetra_train_seq
T_SB1, T_SB2, T_NDB, T_BBK, T_SCH_HU, T_SCH_F = 0, 1, 2, 3, 4, 5          # enum tp_sap_data_type
BLK = {  # type345, type2, type1, a   (tetra_lower_mac.c:55-102)
    T_SB1: (120, 80, 60, 11), T_SB2: (216, 144, 124, 101), T_NDB: (216, 144, 124, 101),
    T_SCH_F: (432, 288, 268, 103), T_BBK: (30, 30, 14, 0), T_SCH_HU: (168, 112, 92, 13),
}

RECORD_DTYPE = np.dtype([
    ("slot_bit", "<u4"), ("lchan", "u1"), ("crc_ok", "u1"), ("blk_num", "u1"),
    ("tn", "u1"), ("fn", "u1"), ("mn", "u1"), ("type1_len", "<u2"),
    ("scrambling_code", "<u4"), ("type1", "u1", (272,)),
])
assert RECORD_DTYPE.itemsize == 288


class GenCfg(C.Structure):
    """struct orc_gen_cfg (oracle/tetra_oracle.c) == struct tb200_gen_cfg (include/tetra_b200.h)"""
    _fields_ = [("seed", C.c_uint64), ("sb_period", C.c_uint32), ("lead_sb", C.c_uint32),
                ("ndb2_per_256", C.c_uint32), ("ber_per_65536", C.c_uint32),
                ("random_cell", C.c_uint32), ("lead_in_bits", C.c_uint32)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


SLOT_DTYPE = np.dtype([("slot_bit", "<u4"), ("scrambling_code", "<u4"), ("find_off", "<u2"),
                       ("window", "<u2"), ("time", "<u2"), ("find_rc", "i1"), ("flags", "u1")])
assert SLOT_DTYPE.itemsize == 16

TB200_FRESH, TB200_FINAL = 1, 2
OUT_UNPACKED, OUT_PACKED = 1, 2
IN_BYTES, IN_PACKED, IN_F32SYM = 0, 1, 2
VITERBI_WARP, VITERBI_LANE = 0, 1
TIE_LOW_PRED, TIE_HIGH_PRED = 0, 1        # include/tetra_tie_rule.h


class Options(C.Structure):
    _fields_ = [("chunk_bits", C.c_uint32), ("output", C.c_uint32), ("viterbi", C.c_uint32),
                ("pipeline_slots", C.c_uint32), ("profile", C.c_uint32), ("input", C.c_uint32),
                ("serial_passes", C.c_uint32), ("afc", C.c_uint32), ("afc_filter_val", C.c_float), ("afc_filter_goal", C.c_float),
                ("viterbi_tie", C.c_uint32), ("host_pack_threads", C.c_uint32)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("classify_ms", C.c_float), ("scan_ms", C.c_float), ("decode_ms", C.c_float),
                ("launches_classify", C.c_uint32), ("launches_scan", C.c_uint32), ("launches_decode", C.c_uint32),
                ("pieces", C.c_uint32), ("slots", C.c_uint64), ("leaf_ms", C.c_float), ("search_ms", C.c_float),
                ("prepare_ms", C.c_float), ("trellis_ms", C.c_float)]


class Carry(C.Structure):
    _fields_ = [("stream_bits", C.c_uint64), ("buf_start_bit", C.c_uint64), ("next_frame_start", C.c_uint64),
                ("calls", C.c_uint64), ("state", C.c_uint32), ("bits_in_buf", C.c_uint32),
                ("scramb_init", C.c_uint32), ("mcc", C.c_uint16), ("mnc", C.c_uint16),
                ("colour_code", C.c_uint8), ("tn", C.c_uint8), ("fn", C.c_uint8), ("mn", C.c_uint8)]


class ShardSummary(C.Structure):
    _fields_ = [("n_slots", C.c_uint32), ("first_unlock", C.c_uint32), ("has_good_sb", C.c_uint32),
                ("scramb_init", C.c_uint32), ("slots_after", C.c_uint32), ("mcc", C.c_uint16), ("mnc", C.c_uint16),
                ("tn", C.c_uint8), ("fn", C.c_uint8), ("mn", C.c_uint8), ("cc", C.c_uint8), ("pad", C.c_uint32)]


assert C.sizeof(ShardSummary) == 32


class DistRun(C.Structure):
    _fields_ = [("global_slot", C.c_uint64), ("local_slot", C.c_uint64), ("n_slots", C.c_uint64)]


class DistTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("total_ms", "pack_ms", "lock_ms", "transfer_ms", "pass1_ms", "exchange_ms", "pass2_ms")] + \
               [("bytes_sent", C.c_uint64), ("segments", C.c_uint32), ("pad", C.c_uint32)]


BCAST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
SCATTER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p, C.c_int)


class DistOps(C.Structure):
    _fields_ = [("user", C.c_void_p), ("bcast", BCAST_FN), ("allgather", ALLGATHER_FN), ("scatter", SCATTER_FN)]


DIST_SCATTER, DIST_PEER, DIST_PACK = 0, 1, 0x100
DIST_ID_BYTES = 128


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("slots", "bursts_decoded", "blocks", "crc_ok_blocks",
                                           "lock_losses", "lock_acquisitions", "kernel_launches")]


class B200:
    """ctypes view of include/tetra_b200.h.  lib_path: another build of the same C ABI (the tests pass the
    SIMT-emulation build of the CUDA sources); by default the in-tree CUDA library, and nothing else: there
    is no CPU implementation behind this class."""

    def __init__(self, device=0, lib_path=None):
        path = lib_path or PRODUCT_SO
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing - run __graft_entry__.build() first (nvcc, sm_100a); there is no CPU fallback")
        lib = self.lib = C.CDLL(path)
        lib.tb200_version.restype = C.c_char_p
        lib.tb200_last_error.restype = C.c_char_p
        lib.tb200_last_error.argtypes = [C.c_void_p]
        lib.tb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        lib.tb200_destroy.argtypes = [C.c_void_p]
        lib.tb200_set_options.argtypes = [C.c_void_p, C.POINTER(Options)]
        for f in (lib.tb200_rx_stream_host, lib.tb200_rx_stream_dev):
            f.restype = C.c_long
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.tb200_max_slots.restype = C.c_uint64
        lib.tb200_max_slots.argtypes = [C.c_uint64]
        lib.tb200_expand_records.restype = C.c_size_t
        lib.tb200_expand_records.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.tb200_get_carry.argtypes = [C.c_void_p, C.POINTER(Carry)]
        lib.tb200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        lib.tb200_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
        lib.tb200_measure_int_peak.restype = C.c_double
        lib.tb200_measure_int_peak.argtypes = [C.c_void_p]
        lib.tb200_find_train_seq.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                             C.c_uint32, C.c_void_p, C.c_void_p]
        lib.tb200_decode_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        lib.tb200_descramble_deinterleave.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                                      C.c_uint32, C.c_uint32, C.c_int]
        if hasattr(lib, "tb200_gsmtap_pack"):
            lib.tb200_gsmtap_pack.restype = C.c_longlong
            lib.tb200_gsmtap_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                              C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
        lib.tb200_gen_stream_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        lib.tb200_find_lock.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.tb200_shard_pass1.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_uint32, C.POINTER(ShardSummary)]
        lib.tb200_shard_carry_in.argtypes = [C.c_void_p, C.c_int, C.POINTER(Carry), C.POINTER(Carry)]
        lib.tb200_shard_carry_in.restype = None
        if hasattr(lib, "tb200_dev_alloc"):          # (A/B runs load older builds of the library)
            lib.tb200_dev_alloc.restype = C.c_void_p
            lib.tb200_dev_alloc.argtypes = [C.c_void_p, C.c_size_t]
            lib.tb200_dev_free.argtypes = [C.c_void_p, C.c_void_p]
            lib.tb200_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            lib.tb200_ipc_import.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
            lib.tb200_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
        lib.tb200_shard_pass2.restype = C.c_long
        lib.tb200_shard_pass2.argtypes = [C.c_void_p, C.POINTER(Carry), C.c_void_p, C.c_void_p, C.c_void_p]
        if hasattr(lib, "tb200_dist_create"):
            lib.tb200_dist_get_id.argtypes = [C.c_void_p]
            lib.tb200_dist_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            lib.tb200_dist_create_with_ops.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.POINTER(DistOps)]
            lib.tb200_dist_destroy.argtypes = [C.c_void_p]
            lib.tb200_dist_last_error.restype = C.c_char_p
            lib.tb200_dist_last_error.argtypes = [C.c_void_p]
            lib.tb200_dist_rx_stream.restype = C.c_long
            lib.tb200_dist_rx_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_uint64, C.POINTER(DistRun), C.c_uint32, C.POINTER(C.c_uint32)]
            lib.tb200_dist_max_local_slots.restype = C.c_uint64
            lib.tb200_dist_max_local_slots.argtypes = [C.c_uint64, C.c_int]
            lib.tb200_dist_get_timing.argtypes = [C.c_void_p, C.POINTER(DistTiming)]
            lib.tb200_slots_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
            lib.tb200_pack_bits_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.tb200_host_alloc.restype = C.c_void_p
        lib.tb200_host_alloc.argtypes = [C.c_size_t]
        lib.tb200_host_free.argtypes = [C.c_void_p]
        self.emulate = lib_path is not None
        h = C.c_void_p()
        rc = lib.tb200_create(C.byref(h), device)
        if rc != 0:
            raise RuntimeError(f"tb200_create failed ({rc}): no CUDA device and no CPU path")
        self.h = h
        self.opt = Options()
        lib.tb200_default_options(C.byref(self.opt))

    def close(self):
        if self.h:
            self.lib.tb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def err(self):
        return self.lib.tb200_last_error(self.h).decode()

    def set_options(self, **kw):
        for k, v in kw.items():
            setattr(self.opt, k, v)
        rc = self.lib.tb200_set_options(self.h, C.byref(self.opt))
        if rc:
            raise ValueError(self.err())

    def rx_stream_host(self, bits, flags=TB200_FRESH | TB200_FINAL, max_slots=None):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        if max_slots is None:
            max_slots = int(self.lib.tb200_max_slots(bits.size)) + 16
        slots = np.zeros(max_slots, dtype=SLOT_DTYPE)
        type1 = np.zeros((max_slots, 288), dtype=np.uint8)
        packed = np.zeros((max_slots, 9), dtype=np.uint32)
        n = self.lib.tb200_rx_stream_host(self.h, _ptr(bits), bits.size, flags, _ptr(slots), _ptr(type1), _ptr(packed), max_slots)
        if n < 0:
            raise RuntimeError(f"tb200_rx_stream_host: {n}: {self.err()}")
        return slots[:n], type1[:n], packed[:n]

    def rx_stream_host_raw(self, buf, n_bits, flags=TB200_FRESH | TB200_FINAL):
        """host call with the stream in the format options.input names (buf: any contiguous numpy array)"""
        buf = np.ascontiguousarray(buf)
        max_slots = int(self.lib.tb200_max_slots(n_bits)) + 16
        slots = np.zeros(max_slots, dtype=SLOT_DTYPE)
        type1 = np.zeros((max_slots, 288), dtype=np.uint8)
        packed = np.zeros((max_slots, 9), dtype=np.uint32)
        n = self.lib.tb200_rx_stream_host(self.h, _ptr(buf), n_bits, flags, _ptr(slots), _ptr(type1), _ptr(packed), max_slots)
        if n < 0:
            raise RuntimeError(f"tb200_rx_stream_host: {n}: {self.err()}")
        return slots[:n], type1[:n], packed[:n]

    def rx_stream_dev_raw(self, buf_ptr, n_bits, flags, slots_ptr, type1_ptr, packed_ptr, max_slots):
        """tb200_rx_stream_dev with raw addresses (device pointers; host addresses under the CPU emulation)"""
        n = self.lib.tb200_rx_stream_dev(self.h, C.c_void_p(buf_ptr), n_bits, flags, C.c_void_p(slots_ptr),
                                         C.c_void_p(type1_ptr) if type1_ptr else None, C.c_void_p(packed_ptr) if packed_ptr else None, max_slots)
        if n < 0:
            raise RuntimeError(f"tb200_rx_stream_dev: {n}: {self.err()}")
        return n

    def expand_records(self, slots, type1):
        slots = np.ascontiguousarray(slots); type1 = np.ascontiguousarray(type1)
        n = self.lib.tb200_expand_records(_ptr(slots), _ptr(type1), slots.size, None, 0)
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        self.lib.tb200_expand_records(_ptr(slots), _ptr(type1), slots.size, _ptr(rec), n)
        return rec

    def records_host(self, bits, flags=TB200_FRESH | TB200_FINAL):
        slots, type1, _ = self.rx_stream_host(bits, flags)
        return self.expand_records(slots, type1)

    def carry(self):
        c = Carry()
        self.lib.tb200_get_carry(self.h, C.byref(c))
        return c

    def stats(self):
        s = Stats()
        self.lib.tb200_get_stats(self.h, C.byref(s))
        return s

    def timing(self):
        t = Timing()
        self.lib.tb200_get_timing(self.h, C.byref(t))
        return t

    # ---- sharded decode (pointers are raw device addresses; host addresses under emulation)
    def find_lock(self, d_bits_ptr, n_bits):
        a0, cmin = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.tb200_find_lock(self.h, C.c_void_p(d_bits_ptr), n_bits, C.byref(a0), C.byref(cmin))
        if rc < 0:
            raise RuntimeError(self.err())
        return rc == 1, a0.value, cmin.value

    def shard_pass1(self, d_bits_ptr, base_bit, n_bytes, a0, cmin, n_end, n_slots):
        s = ShardSummary()
        rc = self.lib.tb200_shard_pass1(self.h, C.c_void_p(d_bits_ptr), base_bit, n_bytes, a0, cmin, n_end, n_slots, C.byref(s))
        if rc:
            raise RuntimeError(self.err())
        return s

    def shard_carry_in(self, summaries, rank, initial=None):
        arr = (ShardSummary * len(summaries))(*summaries)
        init = initial or Carry()
        out = Carry()
        self.lib.tb200_shard_carry_in(arr, rank, C.byref(init), C.byref(out))
        return out

    def shard_pass2(self, carry, d_slots_ptr, d_type1_ptr, d_packed_ptr=None):
        n = self.lib.tb200_shard_pass2(self.h, C.byref(carry), C.c_void_p(d_slots_ptr), C.c_void_p(d_type1_ptr),
                                       C.c_void_p(d_packed_ptr) if d_packed_ptr else None)
        if n < 0:
            raise RuntimeError(self.err())
        return n

    def slots_digest(self, slots, packed, k_base=0, is_device=False, n=None):
        """tb200_slots_digest: numpy arrays (host) or raw device pointers + n (is_device)"""
        out = C.c_uint64(0)
        if is_device:
            r = self.lib.tb200_slots_digest(self.h, C.c_void_p(slots), C.c_void_p(packed) if packed else None, n, k_base, 1, C.byref(out))
        else:
            slots = np.ascontiguousarray(slots); packed = np.ascontiguousarray(packed, dtype=np.uint32)
            r = self.lib.tb200_slots_digest(self.h, _ptr(slots), _ptr(packed), slots.size, k_base, 0, C.byref(out))
        if r:
            raise RuntimeError(self.err())
        return out.value

    def viterbi_decode(self, mother, sym_count):
        """tb200_viterbi_decode over the rows of `mother` (n x 4*sym_count bytes: 0 / 1 / 0xff) -> n x sym_count bits"""
        mother = np.ascontiguousarray(mother, dtype=np.uint8).reshape(-1, 4 * sym_count)
        out = np.zeros((mother.shape[0], sym_count), dtype=np.uint8)
        self.lib.tb200_viterbi_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        if self.lib.tb200_viterbi_decode(self.h, _ptr(mother), mother.shape[0], sym_count, _ptr(out), 0):
            raise RuntimeError(self.err())
        return out

    def float_to_bits(self, sym, afc=False, filter_val=0.0001, filter_goal=0.0, state=0.0):
        """tb200_float_to_bits on a host array -> (unpacked bits, tracker state afterwards, chunks redone)"""
        sym = np.ascontiguousarray(sym, dtype=np.float32)
        out = np.zeros(4 * ((sym.size + 15) // 16) + 4, dtype=np.uint8)
        st = C.c_float(state)
        self.lib.tb200_float_to_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_float,
                                                 C.POINTER(C.c_float), C.c_void_p, C.c_int]
        rc = self.lib.tb200_float_to_bits(self.h, _ptr(sym), sym.size, int(afc), filter_val, filter_goal, C.byref(st), _ptr(out), 0)
        if rc < 0:
            raise RuntimeError(self.err())
        return np.unpackbits(out, bitorder="little")[:2 * sym.size], st.value, rc

    def rm3014_decode(self, words):
        """tb200_rm3014_decode on a host array of 30-bit words -> (info14, distance, not_a_code_word)"""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        out = np.zeros(words.size, dtype=np.uint32)
        self.lib.tb200_rm3014_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        if self.lib.tb200_rm3014_decode(self.h, _ptr(words), words.size, _ptr(out), 0):
            raise RuntimeError(self.err())
        return out & 0x3fff, (out >> 16) & 0xff, (out >> 24) & 1

    def rx_stream_host_aach(self, bits):
        """rx_stream_host with the RM(30,14)-decoded AACH side output switched on for this call"""
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        ms = int(self.lib.tb200_max_slots(bits.size)) + 16
        aach = np.zeros(ms, dtype=np.uint32)
        self.lib.tb200_set_aach_buffer.argtypes = [C.c_void_p, C.c_void_p]
        if self.lib.tb200_set_aach_buffer(self.h, _ptr(aach)):
            raise RuntimeError(self.err())
        try:
            slots, t1, pk = self.rx_stream_host(bits, max_slots=ms)
        finally:
            self.lib.tb200_set_aach_buffer(self.h, None)
        return slots, t1, pk, aach[:slots.size]

    def find_train_seq(self, bits, starts, lens, mask):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64); lens = np.ascontiguousarray(lens, dtype=np.uint32)
        rc = np.zeros(starts.size, dtype=np.int32); off = np.zeros(starts.size, dtype=np.uint32)
        r = self.lib.tb200_find_train_seq(self.h, _ptr(bits), bits.size, _ptr(starts), _ptr(lens), starts.size, mask, _ptr(rc), _ptr(off))
        if r:
            raise RuntimeError(self.err())
        return rc, off

    def decode_blocks(self, blk_type, type5, codes):
        k, _, t1, _ = BLK[blk_type]
        type5 = np.ascontiguousarray(type5, dtype=np.uint8).reshape(-1, k)
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        n = type5.shape[0]
        out = np.zeros((n, t1), dtype=np.uint8); ok = np.zeros(n, dtype=np.uint8)
        r = self.lib.tb200_decode_blocks(self.h, blk_type, _ptr(type5), _ptr(codes), n, _ptr(out), _ptr(ok))
        if r:
            raise RuntimeError(self.err())
        return out, ok

    def descramble_deinterleave(self, type5, codes, K, a):
        type5 = np.ascontiguousarray(type5, dtype=np.uint8).reshape(-1, K)
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        out = np.zeros_like(type5)
        r = self.lib.tb200_descramble_deinterleave(self.h, _ptr(type5), _ptr(out), _ptr(codes), type5.shape[0], K, a, 0)
        if r:
            raise RuntimeError(self.err())
        return out

    def rcpc_depunct(self, puncturer, type3, mother_len):
        """tetra_rcpc_depunct over the rows of type3 (n x len bytes) -> n x mother_len bytes, 0xff where nothing was sent"""
        type3 = np.ascontiguousarray(type3, dtype=np.uint8)
        n, ln = type3.shape
        out = np.zeros((n, mother_len), dtype=np.uint8)
        self.lib.tb200_rcpc_depunct.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int]
        r = self.lib.tb200_rcpc_depunct(self.h, puncturer, _ptr(type3), ln, n, _ptr(out), mother_len, 0)
        if r:
            raise RuntimeError(self.err())
        return out

    def gsmtap_pack(self, slots, packed):
        """GSMTAP frames of every CRC-good block of the slots (host arrays): (frame bytes, slot offsets, n_frames)"""
        slots = np.ascontiguousarray(slots); packed = np.ascontiguousarray(packed, dtype=np.uint32)
        nf = C.c_uint64(0)
        need = self.lib.tb200_gsmtap_pack(self.h, _ptr(slots), _ptr(packed), slots.size, None, 0, None, C.byref(nf), 0)
        if need < 0:
            raise RuntimeError(self.err())
        frames = np.zeros(max(need, 2), dtype=np.uint8)
        off = np.zeros(slots.size + 1, dtype=np.uint64)
        got = self.lib.tb200_gsmtap_pack(self.h, _ptr(slots), _ptr(packed), slots.size, _ptr(frames), need, _ptr(off), C.byref(nf), 0)
        if got != need:
            raise RuntimeError(f"tb200_gsmtap_pack: {got} != {need}: {self.err()}")
        return frames[:need], off, nf.value


def slots_digest_host(slots, packed, k_base=0):
    """numpy twin of tb200_slots_digest (csrc/tetra_util.cuh: slot_digest), for checking the device digest"""
    M = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        k = np.arange(slots.size, dtype=np.uint64) + np.uint64(k_base)
        h = (k + np.uint64(1)) * np.uint64(0xD6E8FEB86659FD93)
        h ^= h >> np.uint64(29)

        def mix(h, w):
            h = (h ^ w.astype(np.uint64)) * M
            return h ^ (h >> np.uint64(32))
        h = mix(h, slots["slot_bit"])
        h = mix(h, slots["scrambling_code"])
        h = mix(h, slots["find_off"].astype(np.uint32) | (slots["window"].astype(np.uint32) << np.uint32(16)))
        h = mix(h, slots["time"].astype(np.uint32) | (slots["find_rc"].view(np.uint8).astype(np.uint32) << np.uint32(16)) |
                (slots["flags"].astype(np.uint32) << np.uint32(24)))
        for i in range(9):
            h = mix(h, packed[:, i])
        return int(h.sum(dtype=np.uint64))


class Dist:
    """tb200_dist_*: one stream decoded by `world` ranks.  nccl_id: 128 bytes from Dist.get_id() on rank 0 (handed to the
    other ranks by the caller); or ops: a DistOps with the caller's own plumbing (the CPU tests use gloo)."""

    def __init__(self, g, rank, world, nccl_id=None, ops=None):
        self.g, self.rank, self.world = g, rank, world
        self.h = C.c_void_p()
        self._ops = ops
        if ops is not None:
            rc = g.lib.tb200_dist_create_with_ops(C.byref(self.h), g.h, rank, world, C.byref(ops))
        else:
            buf = (C.c_uint8 * DIST_ID_BYTES).from_buffer_copy(nccl_id)
            rc = g.lib.tb200_dist_create(C.byref(self.h), g.h, rank, world, buf)
        if rc:
            raise RuntimeError(f"tb200_dist_create failed ({rc}): {g.err()}")

    @staticmethod
    def get_id(g):
        buf = (C.c_uint8 * DIST_ID_BYTES)()
        if g.lib.tb200_dist_get_id(buf):
            raise RuntimeError("tb200_dist_get_id failed (libnccl.so.2 missing?)")
        return bytes(buf)

    def err(self):
        return self.g.lib.tb200_dist_last_error(self.h).decode()

    def max_local_slots(self, n_bits):
        return int(self.g.lib.tb200_dist_max_local_slots(n_bits, self.world))

    def rx_stream(self, bits_ptr, n_bits, mode, slots_ptr, type1_ptr, packed_ptr, max_slots, max_runs=64):
        runs = (DistRun * max_runs)()
        nr = C.c_uint32(0)
        n = self.g.lib.tb200_dist_rx_stream(self.h, C.c_void_p(bits_ptr) if bits_ptr else None, n_bits, mode, C.c_void_p(slots_ptr),
                                            C.c_void_p(type1_ptr) if type1_ptr else None, C.c_void_p(packed_ptr) if packed_ptr else None,
                                            max_slots, runs, max_runs, C.byref(nr))
        if n < 0:
            raise RuntimeError(f"tb200_dist_rx_stream: {n}: {self.err()}")
        return n, [(r.global_slot, r.local_slot, r.n_slots) for r in runs[:nr.value]]

    def timing(self):
        t = DistTiming()
        self.g.lib.tb200_dist_get_timing(self.h, C.byref(t))
        return t

    def close(self):
        if self.h:
            self.g.lib.tb200_dist_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------- golden fixtures

def shard_plan(n_slots_total, world):
    """contiguous slot ranges per rank (the last ranks get the remainder)"""
    per = (n_slots_total + world - 1) // world
    return [(min(r * per, n_slots_total), min((r + 1) * per, n_slots_total)) for r in range(world)]


SHARD_HALO = 4096 + 64        # look-ahead of the search window (tetra_burst_sync.c:117) + read-ahead


class DevBuffer:
    """device memory from tb200_dev_alloc (exportable to other processes), viewable as a torch uint8 tensor"""

    def __init__(self, g, nbytes):
        self.g, self.nbytes = g, nbytes
        self.ptr = g.lib.tb200_dev_alloc(g.h, nbytes)
        if not self.ptr:
            raise MemoryError("tb200_dev_alloc")
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=device)

    def export(self):
        h = (C.c_uint8 * 64)()
        if self.g.lib.tb200_ipc_export(self.g.h, C.c_void_p(self.ptr), h):
            raise RuntimeError(self.g.err())
        return bytes(h)

    def free(self):
        if self.ptr:
            self.g.lib.tb200_dev_free(self.g.h, C.c_void_p(self.ptr))
            self.ptr = None


_peer_maps = {}


def peer_pointer(g, handle):
    """map another rank's exported buffer (cached per handle: opening is a millisecond-scale call)"""
    key = (id(g), handle)
    if key not in _peer_maps:
        p = C.c_void_p()
        hb = (C.c_uint8 * 64).from_buffer_copy(handle)
        if g.lib.tb200_ipc_import(g.h, hb, C.byref(p)):
            raise RuntimeError(g.err())
        _peer_maps[key] = p.value
    return _peer_maps[key]


def sharded_decode(g, dist, rank, world, d_full, n_bits, device, want_type1=False, timers=None, peer_handle=None):
    """BASELINE config 5: ONE stream, held by rank 0 (device tensor d_full, None elsewhere), decoded by
    `world` ranks.  Rank 0 acquires lock on the head of the stream; every rank takes a contiguous slot range
    (+ look-ahead halo) and runs pass 1 (search, classification, SB1) on it, the 32-byte summaries are
    all-gathered (the only exchange step of the path: the cell state), every rank derives its carry-in and
    runs pass 2.  Results stay rank-local.  How a rank gets at its shard:
      peer_handle is None   rank 0 scatters the shards with NCCL send/recv, then the kernels run
      peer_handle = bytes   (rank 0's exported buffer, see DevBuffer) no copy at all: the search kernel of
                            every rank reads its shard straight out of rank 0's HBM over NVLink
    Returns (k0, k1, a0, d_slots, d_type1 or None, d_packed, summaries)."""
    import torch
    meta = torch.zeros(4, dtype=torch.int64, device=device)
    if rank == 0:
        ok, a0, cmin = g.find_lock(d_full.data_ptr(), n_bits)
        meta[:] = torch.tensor([int(ok), a0, cmin, n_bits], dtype=torch.int64)
    if world > 1:
        dist.broadcast(meta, 0)
    ok, a0, cmin, n_bits = (int(x) for x in meta.cpu())
    if not ok:
        raise RuntimeError("no lock on the head of the stream")
    n_total = (n_bits - a0) // 510
    plan = shard_plan(n_total, world)
    k0, k1 = plan[rank]
    n = k1 - k0

    def span(r):
        lo = a0 + 510 * plan[r][0]
        hi = min(n_bits, a0 + 510 * plan[r][1] + SHARD_HALO)
        return lo, max(hi, lo)
    if timers is not None:
        torch.cuda.synchronize(); timers["t_scatter0"] = __import__("time").perf_counter()
    lo, hi = span(rank)
    if peer_handle is not None:
        shard_ptr = (d_full.data_ptr() if rank == 0 else peer_pointer(g, peer_handle)) + lo
    elif rank == 0:
        shard = d_full[lo:hi]
        shard_ptr = shard.data_ptr()
        if world > 1:
            # one NCCL group: the sends to all peers run concurrently and share rank 0's NVLink egress
            ops = [dist.P2POp(dist.isend, d_full[span(r)[0]:span(r)[1]], r) for r in range(1, world) if span(r)[1] > span(r)[0]]
            for q in (dist.batch_isend_irecv(ops) if ops else []):
                q.wait()
    else:
        shard = torch.empty(hi - lo + 64, dtype=torch.uint8, device=device)[:hi - lo]
        shard_ptr = shard.data_ptr()
        if hi > lo:
            for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, shard, 0)]):
                q.wait()
    if timers is not None:
        if world > 1:
            dist.barrier()          # the scatter ends when the last rank has its shard
        torch.cuda.synchronize(); timers["t_scatter1"] = __import__("time").perf_counter()
    s = g.shard_pass1(shard_ptr, lo, hi - lo, lo, cmin + k0, n_bits, n)
    mine = torch.frombuffer(bytearray(bytes(s)), dtype=torch.uint8).to(device)
    gathered = [torch.zeros(32, dtype=torch.uint8, device=device) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, mine)
    else:
        gathered[0] = mine
    summaries = [ShardSummary.from_buffer_copy(bytes(t.cpu().numpy().tobytes())) for t in gathered]
    carry = g.shard_carry_in(summaries, rank)
    d_slots = torch.empty(max(n, 1) * 16, dtype=torch.uint8, device=device)
    d_t1 = torch.empty(max(n, 1) * 288, dtype=torch.uint8, device=device) if want_type1 else None
    d_pk = torch.empty(max(n, 1) * 9, dtype=torch.int32, device=device)
    got = g.shard_pass2(carry, d_slots.data_ptr(), d_t1.data_ptr() if want_type1 else None, d_pk.data_ptr())
    assert got == n, (got, n)
    if timers is not None:
        torch.cuda.synchronize(); timers["t_done"] = __import__("time").perf_counter()
    return k0, k1, a0, d_slots, d_t1, d_pk, summaries


def pack_bits(bits):
    """1-bit-per-byte stream -> TB200_IN_PACKED buffer (stream bit i = byte i>>3 bit i&7), padded to 16 bytes"""
    p = np.packbits(np.ascontiguousarray(bits, dtype=np.uint8) & 1, bitorder="little")
    out = np.zeros((p.size + 15) // 16 * 16, dtype=np.uint8)
    out[:p.size] = p
    return out

