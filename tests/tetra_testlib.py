"""ctypes bindings used by the tests (and by bench.py / smoke() for the checker legs).

Three libraries, all loaded lazily:
  * oracle/_ref/libtetra_ref.so   - the reference's own lower MAC compiled in place
                                    (oracle/Makefile `make ref`); prebuilt, travels
                                    to the GPU box, never rebuilt there.
  * oracle/libtetra_oracle.so     - the stand-alone CPU restatement (`make oracle`).
  * osmo-tetra_b200/libtetra_b200.so - the product (CUDA + C-ABI), see include/tetra_b200.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libtetra_ref.so")
ORACLE_SO = os.path.join(ORACLE_DIR, "libtetra_oracle.so")

# enum tetra_train_seq, tetra_burst.h:27-33
TS_NORM_1, TS_NORM_2, TS_NORM_3, TS_SYNC, TS_EXT = 0, 1, 2, 3, 4
# enum tp_sap_data_type, tetra_burst.h:9-16
T_SB1, T_SB2, T_NDB, T_BBK, T_SCH_HU, T_SCH_F = 0, 1, 2, 3, 4, 5
# enum tetra_log_chan, tetra_common.h:22-39
LC_UNKNOWN, LC_SCH_F, LC_AACH, LC_BSCH, LC_BNCH = 0, 1, 8, 10, 11

BLK = {  # type345, type2, type1, a   (tetra_lower_mac.c:55-102)
    T_SB1: (120, 80, 60, 11),
    T_SB2: (216, 144, 124, 101),
    T_NDB: (216, 144, 124, 101),
    T_BBK: (30, 30, 14, 0),
    T_SCH_HU: (168, 112, 92, 13),
    T_SCH_F: (432, 288, 268, 103),
}

EVENT_DTYPE = np.dtype([
    ("call_index", "<u4"), ("buf_start_bit", "<u4"), ("window", "<u4"),
    ("mask", "<u4"), ("rc", "<i4"), ("offset", "<u4"),
])
assert EVENT_DTYPE.itemsize == 24


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def bits_from_str(s):
    return np.frombuffer(s.encode(), dtype=np.uint8) - ord("0")


def bits_to_str(a):
    return "".join(str(int(x)) for x in a)


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "tetra_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    if os.path.isdir("/root/reference/src") and not os.path.exists(REF_SO):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


class _Recorder:
    """common record/event access for the two CPU libraries (prefix ref_ / orc_)"""

    def __init__(self, lib, prefix):
        self.lib, self.p = lib, prefix
        f = lambda n: getattr(lib, prefix + n)
        f("num_records").restype = C.c_size_t
        f("records").restype = C.c_void_p
        f("num_events").restype = C.c_size_t
        f("events").restype = C.c_void_p
        f("cell_scramb_init").restype = C.c_uint32
        self._f = f

    def reset(self):
        self._f("reset")()

    def records(self):
        n = self._f("num_records")()
        if n == 0:
            return np.zeros(0, dtype=RECORD_DTYPE)
        buf = C.string_at(self._f("records")(), n * RECORD_DTYPE.itemsize)
        return np.frombuffer(buf, dtype=RECORD_DTYPE).copy()

    def events(self):
        n = self._f("num_events")()
        if n == 0:
            return np.zeros(0, dtype=EVENT_DTYPE)
        buf = C.string_at(self._f("events")(), n * EVENT_DTYPE.itemsize)
        return np.frombuffer(buf, dtype=EVENT_DTYPE).copy()

    def rx_state(self):
        return self._f("rx_state")()

    def scramb_init(self):
        return self._f("cell_scramb_init")()

    def set_recording(self, on):
        self._f("set_recording")(int(on))

    def gsmtap_frames(self, records):
        """GSMTAP frames of the CRC-good records, back to back (bytes), and their number"""
        f = self._f("gsmtap_frames")
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        records = np.ascontiguousarray(records)
        nf = C.c_size_t(0)
        need = f(_ptr(records), records.size, None, 0, C.byref(nf))
        out = np.zeros(max(need, 1), dtype=np.uint8)
        f(_ptr(records), records.size, _ptr(out), need, C.byref(nf))
        return out[:need], nf.value

    def set_time(self, tn, fn, mn):
        self._f("set_time")(C.c_uint32(tn), C.c_uint32(fn), C.c_uint32(mn))

    def get_time(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._f("get_time")(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class Ref(_Recorder):
    """The reference's own code (oracle/_ref)."""

    def __init__(self):
        lib = C.CDLL(REF_SO)
        super().__init__(lib, "ref_")
        lib.tetra_scramb_get_init.restype = C.c_uint32
        lib.tetra_rm3014_compute.restype = C.c_uint32
        lib.crc16_ccitt_bits.restype = C.c_uint16
        lib.ref_feed.restype = C.c_long
        lib.ref_feed.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_int]
        lib.ref_rm3014_init()

    def feed(self, bits, chunk=64):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        return self.lib.ref_feed(_ptr(bits), bits.size, chunk, 1)

    def tp_sap(self, typ, blk_num, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        self.lib.ref_tp_sap(typ, blk_num, _ptr(bits), bits.size, 1)

    def set_tie(self, tie):
        """tie rule of the osmo_conv_decode stand-in the reference code is linked against (libosmocore is absent)"""
        self.lib.oracle_conv_set_tie(int(tie))

    def find_train_seq(self, window, end, mask):
        """window must have >= end+21 readable bytes"""
        off = C.c_uint(0)
        w = np.ascontiguousarray(window, dtype=np.uint8)
        rc = self.lib.__real_tetra_find_train_seq(_ptr(w), C.c_uint(end), C.c_uint32(mask), C.byref(off)) \
            if hasattr(self.lib, "__real_tetra_find_train_seq") else \
            self.lib.tetra_find_train_seq(_ptr(w), C.c_uint(end), C.c_uint32(mask), C.byref(off))
        return rc, off.value

    def scramb_get_bits(self, init, n):
        out = np.zeros(n, dtype=np.uint8)
        self.lib.tetra_scramb_get_bits(C.c_uint32(init), _ptr(out), n)
        return out

    def scramb_bits(self, init, bits):
        out = np.array(bits, dtype=np.uint8)
        self.lib.tetra_scramb_bits(C.c_uint32(init), _ptr(out), out.size)
        return out

    def scramb_get_init(self, mcc, mnc, cc):
        return self.lib.tetra_scramb_get_init(C.c_uint16(mcc), C.c_uint16(mnc), C.c_uint8(cc))

    def deinterleave(self, K, a, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8); o = np.zeros(K, dtype=np.uint8)
        self.lib.block_deinterleave(C.c_uint32(K), C.c_uint32(a), _ptr(i), _ptr(o))
        return o

    def interleave(self, K, a, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8); o = np.zeros(K, dtype=np.uint8)
        self.lib.block_interleave(C.c_uint32(K), C.c_uint32(a), _ptr(i), _ptr(o))
        return o

    def depunct_2_3(self, type3, n_mother):
        i = np.ascontiguousarray(type3, dtype=np.uint8); o = np.full(n_mother, 0xff, dtype=np.uint8)
        self.lib.tetra_rcpc_depunct(0, _ptr(i), i.size, _ptr(o))
        return o

    def punct_2_3(self, mother, n_type3):
        i = np.ascontiguousarray(mother, dtype=np.uint8); o = np.zeros(n_type3, dtype=np.uint8)
        self.lib.get_punctured_rate(0, _ptr(i), n_type3, _ptr(o))
        return o

    def conv_encode(self, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8); o = np.zeros(4 * i.size, dtype=np.uint8)
        st = (C.c_uint8 * 4)()
        self.lib.conv_enc_init(st)
        self.lib.conv_enc_input(st, _ptr(i), i.size, _ptr(o))
        return o

    def viterbi(self, mother, n):
        i = np.ascontiguousarray(mother, dtype=np.uint8); o = np.zeros(n, dtype=np.uint8)
        self.lib.viterbi_dec_sb1_wrapper(_ptr(i), _ptr(o), C.c_uint(n))
        return o

    def crc16(self, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8)
        return self.lib.crc16_ccitt_bits(_ptr(i), C.c_uint(i.size))

    def rm3014(self, info):
        return self.lib.tetra_rm3014_compute(C.c_uint16(info))

    def build_sync_burst(self, sb, bb, bkn):
        out = np.zeros(510, dtype=np.uint8)
        a, b, c = (np.ascontiguousarray(x, dtype=np.uint8) for x in (sb, bb, bkn))
        n = self.lib.ref_build_sync_burst(_ptr(out), _ptr(a), _ptr(b), _ptr(c))
        assert n == 510
        return out

    def build_norm_burst(self, bkn1, bb, bkn2, two):
        out = np.zeros(510, dtype=np.uint8)
        a, b, c = (np.ascontiguousarray(x, dtype=np.uint8) for x in (bkn1, bb, bkn2))
        n = self.lib.ref_build_norm_burst(_ptr(out), _ptr(a), _ptr(b), _ptr(c), int(two))
        assert n == 510
        return out

    def time_add_slot(self, tn, fn, mn):
        class T(C.Structure):
            _fields_ = [("hn", C.c_uint16), ("sn", C.c_uint32), ("tn", C.c_uint32),
                        ("fn", C.c_uint32), ("mn", C.c_uint32)]
        t = T(0, 0, tn, fn, mn)
        self.lib.tetra_tdma_time_add_tn(C.byref(t), C.c_uint32(1))
        return t.tn, t.fn, t.mn


class Oracle(_Recorder):
    """The stand-alone CPU restatement (oracle/tetra_oracle.c)."""

    def __init__(self):
        ensure_oracle_built()
        lib = C.CDLL(ORACLE_SO)
        super().__init__(lib, "orc_")
        lib.orc_scramb_get_init.restype = C.c_uint32
        lib.orc_rm3014_compute.restype = C.c_uint32
        lib.orc_crc16.restype = C.c_uint16
        lib.orc_feed.restype = C.c_long
        lib.orc_feed.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
        lib.orc_records_digest.restype = C.c_uint64
        lib.orc_records_digest.argtypes = [C.c_void_p, C.c_size_t]
        lib.orc_gen_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        lib.orc_gen_burst.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_gen_kind.argtypes = [C.c_void_p, C.c_uint64]
        lib.orc_float_to_bits.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]

    def feed(self, bits, chunk=64):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        return self.lib.orc_feed(_ptr(bits), bits.size, chunk)

    def float_to_bits(self, sym):
        """float_to_bits.c without AFC: float32 symbols -> 2 unpacked bits each"""
        sym = np.ascontiguousarray(sym, dtype=np.float32)
        out = np.zeros(2 * sym.size, dtype=np.uint8)
        self.lib.orc_float_to_bits(_ptr(sym), sym.size, _ptr(out))
        return out

    def float_to_bits_afc(self, sym, filter_val=0.0001, filter_goal=0.0, state=0.0):
        """float_to_bits -a [-f filter_val] [-F filter_goal]: (bits, tracker state afterwards)"""
        sym = np.ascontiguousarray(sym, dtype=np.float32)
        out = np.zeros(2 * sym.size, dtype=np.uint8)
        f = C.c_float(state)
        self.lib.orc_float_to_bits_afc.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.POINTER(C.c_float), C.c_void_p]
        self.lib.orc_float_to_bits_afc(_ptr(sym), sym.size, C.c_float(filter_val), C.c_float(filter_goal), C.byref(f), _ptr(out))
        return out, f.value

    def tp_sap(self, typ, blk_num, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        self.lib.orc_tp_sap(typ, blk_num, _ptr(bits))

    def set_cell(self, scramb_init):
        self.lib.orc_set_cell(C.c_uint32(scramb_init))

    def set_tie(self, tie):
        """Viterbi tie rule of the restatement (include/tetra_tie_rule.h); the default is the compile-time switch"""
        self.lib.orc_set_tie(int(tie))

    def find_train_seq(self, window, end, mask):
        off = C.c_uint(0)
        w = np.ascontiguousarray(window, dtype=np.uint8)
        rc = self.lib.orc_find_train_seq(_ptr(w), C.c_uint(end), C.c_uint32(mask), C.byref(off))
        return rc, off.value

    def scramb_get_bits(self, init, n):
        out = np.zeros(n, dtype=np.uint8)
        self.lib.orc_scramb_get_bits(C.c_uint32(init), _ptr(out), n)
        return out

    def scramb_bits(self, init, bits):
        out = np.array(bits, dtype=np.uint8)
        self.lib.orc_scramb_bits(C.c_uint32(init), _ptr(out), out.size)
        return out

    def scramb_get_init(self, mcc, mnc, cc):
        return self.lib.orc_scramb_get_init(mcc, mnc, cc)

    def deinterleave(self, K, a, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8); o = np.zeros(K, dtype=np.uint8)
        self.lib.orc_deinterleave(K, a, _ptr(i), _ptr(o))
        return o

    def interleave(self, K, a, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8); o = np.zeros(K, dtype=np.uint8)
        self.lib.orc_interleave(K, a, _ptr(i), _ptr(o))
        return o

    def depunct_2_3(self, type3, n_mother):
        i = np.ascontiguousarray(type3, dtype=np.uint8); o = np.full(n_mother, 0xff, dtype=np.uint8)
        self.lib.orc_depunct_2_3(_ptr(i), i.size, _ptr(o))
        return o

    def punct_2_3(self, mother, n_type3):
        i = np.ascontiguousarray(mother, dtype=np.uint8); o = np.zeros(n_type3, dtype=np.uint8)
        self.lib.orc_punct_2_3(_ptr(i), n_type3, _ptr(o))
        return o

    def conv_encode(self, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8); o = np.zeros(4 * i.size, dtype=np.uint8)
        self.lib.orc_conv_encode(_ptr(i), i.size, _ptr(o))
        return o

    def viterbi(self, mother, n):
        i = np.ascontiguousarray(mother, dtype=np.uint8); o = np.zeros(n, dtype=np.uint8)
        self.lib.orc_viterbi(_ptr(i), _ptr(o), n)
        return o

    def crc16(self, bits):
        i = np.ascontiguousarray(bits, dtype=np.uint8)
        return self.lib.orc_crc16(_ptr(i), i.size)

    def rm3014(self, info):
        return self.lib.orc_rm3014_compute(C.c_uint16(info))

    def rm3014_decode_ml(self, word):
        """brute-force nearest code word -> (info14, distance)"""
        info = C.c_uint16(0)
        d = self.lib.orc_rm3014_decode_ml(C.c_uint32(int(word)), C.byref(info))
        return info.value, d

    def time_add_slot(self, tn, fn, mn):
        t = (C.c_uint32 * 3)(tn, fn, mn)
        self.lib.orc_time_add_slot(t)
        return t[0], t[1], t[2]

    def digest(self, records):
        r = np.ascontiguousarray(records)
        return self.lib.orc_records_digest(_ptr(r), r.size)

    # ---- generator
    def gen_stream(self, cfg, k0, n, lead_in=True):
        nbits = 510 * n + (cfg.lead_in_bits if lead_in else 0)
        out = np.zeros(nbits, dtype=np.uint8)
        self.lib.orc_gen_stream(C.byref(cfg), k0, n, _ptr(out), int(lead_in))
        return out

    def gen_kind(self, cfg, k):
        return self.lib.orc_gen_kind(C.byref(cfg), k)


def have_ref():
    return os.path.exists(REF_SO)


def records_equal(a, b):
    """bit-exact comparison of two record arrays; returns (ok, message)"""
    if a.size != b.size:
        return False, f"record count {a.size} != {b.size}"
    for name in RECORD_DTYPE.names:
        if not np.array_equal(a[name], b[name]):
            bad = np.nonzero(np.any(a[name] != b[name], axis=-1) if a[name].ndim > 1 else a[name] != b[name])[0]
            return False, f"field {name} differs at records {bad[:5]} (of {bad.size})"
    return True, ""


# --------------------------------------------------------------------------- product

PRODUCT_SO = os.path.join(ROOT, "osmo-tetra_b200", "libtetra_b200.so")
SIMT_SO = os.path.join(ROOT, "tests", "simt", "_build", "libtetra_b200_simt.so")

# ---- the product's own ctypes binding (osmo-tetra_b200/binding.py), re-exported for the tests

def _load_binding():
    import importlib.util
    import sys
    name = "osmo_tetra_b200_binding"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "osmo-tetra_b200", "binding.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_B = _load_binding()
RECORD_DTYPE, SLOT_DTYPE, GenCfg, Options, Timing, Carry, ShardSummary, Stats = (
    _B.RECORD_DTYPE, _B.SLOT_DTYPE, _B.GenCfg, _B.Options, _B.Timing, _B.Carry, _B.ShardSummary, _B.Stats)
TB200_FRESH, TB200_FINAL, OUT_UNPACKED, OUT_PACKED = _B.TB200_FRESH, _B.TB200_FINAL, _B.OUT_UNPACKED, _B.OUT_PACKED
IN_BYTES, IN_PACKED, IN_F32SYM, VITERBI_WARP, VITERBI_LANE = _B.IN_BYTES, _B.IN_PACKED, _B.IN_F32SYM, _B.VITERBI_WARP, _B.VITERBI_LANE
TIE_LOW_PRED, TIE_HIGH_PRED = _B.TIE_LOW_PRED, _B.TIE_HIGH_PRED
shard_plan, SHARD_HALO, DevBuffer, peer_pointer, sharded_decode, pack_bits = (
    _B.shard_plan, _B.SHARD_HALO, _B.DevBuffer, _B.peer_pointer, _B.sharded_decode, _B.pack_bits)


Dist, DistOps, DIST_SCATTER, DIST_PEER, DIST_PACK, slots_digest_host = (
    _B.Dist, _B.DistOps, _B.DIST_SCATTER, _B.DIST_PEER, _B.DIST_PACK, _B.slots_digest_host)


def gloo_dist_ops(dist):
    """tb200_dist_ops over torch.distributed (gloo, host memory): the plumbing the CPU tests hand to the C driver
    (under the SIMT emulation "device" pointers are host pointers).  Keep the returned object alive."""
    import torch

    def view(ptr, n):
        return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)))

    def bcast(user, buf, n, root):
        dist.broadcast(view(buf, n), root)
        return 0

    def allgather(user, send, recv, each):
        world = dist.get_world_size()
        out = [torch.zeros(each, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(out, view(send, each).clone())
        view(recv, each * world).copy_(torch.cat(out))
        return 0

    def scatter(user, src, offs, sizes, dst, root):
        rank, world = dist.get_rank(), dist.get_world_size()
        if rank == root:
            for r in range(world):
                if r != root:
                    dist.send(torch.tensor([sizes[r]], dtype=torch.int64), r)
                    if sizes[r]:
                        dist.send(view(src + offs[r], sizes[r]).clone(), r)
        else:
            n = torch.zeros(1, dtype=torch.int64)
            dist.recv(n, root)
            if int(n[0]) != sizes[rank]:          # sender and receiver must agree on the shard's size on the wire
                return -2
            if sizes[rank]:
                t = torch.zeros(sizes[rank], dtype=torch.uint8)
                dist.recv(t, root)
                view(dst, sizes[rank]).copy_(t)
        return 0

    ops = _B.DistOps(None, _B.BCAST_FN(bcast), _B.ALLGATHER_FN(allgather), _B.SCATTER_FN(scatter))
    return ops


def B200(emulate=False, device=0):
    """the product binding; emulate=True: on the SIMT-emulation build of the CUDA sources (CPU tests)"""
    return _B.B200(device=device, lib_path=build_simt() if emulate else None)


def build_simt():
    """(re)build the CPU SIMT-emulation flavour of the library - test infrastructure only"""
    csrc = os.path.join(ROOT, "osmo-tetra_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + \
           [os.path.join(ROOT, "tests", "simt", f) for f in ("cpu_simt.h", "cpu_simt.cpp")] + \
           [os.path.join(ROOT, "include", "tetra_b200.h")]
    if os.path.exists(SIMT_SO) and all(os.path.getmtime(SIMT_SO) >= os.path.getmtime(s) for s in srcs):
        return SIMT_SO
    os.makedirs(os.path.dirname(SIMT_SO), exist_ok=True)
    subprocess.check_call([
        "g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-x", "c++",
        "-include", os.path.join(ROOT, "tests", "simt", "cpu_simt.h"),
        "-I" + os.path.join(ROOT, "osmo-tetra_b200", "csrc"),
        os.path.join(ROOT, "osmo-tetra_b200", "csrc", "tetra_b200.cu"),
        os.path.join(ROOT, "tests", "simt", "cpu_simt.cpp"), "-o", SIMT_SO])
    return SIMT_SO


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def load_golden_stream(name):
    """-> (bits, records, events) of a fixture written by tests/golden/make_golden.py"""
    z = np.load(os.path.join(GOLDEN_DIR, name))
    bits = np.unpackbits(z["bits"])[:int(z["n_bits"])]
    n = z["slot_bit"].size
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    for f in ("slot_bit", "lchan", "crc_ok", "blk_num", "tn", "fn", "mn", "type1_len", "scrambling_code"):
        rec[f] = z[f]
    rec["type1"] = np.unpackbits(z["type1"], axis=1)[:, :272]
    return bits, rec, z["events"]


def load_golden_blocks(name):
    z = np.load(os.path.join(GOLDEN_DIR, "blocks_noisy.npz"))
    bt = {"sb1": T_SB1, "ndb": T_NDB, "schf": T_SCH_F}[name]
    K, _, T1, _ = BLK[bt]
    return (bt, np.unpackbits(z[name + "_type5"], axis=1)[:, :K], z[name + "_code"],
            np.unpackbits(z[name + "_type1"], axis=1)[:, :T1], z[name + "_crc_ok"])


def locked_events(events):
    """the searches the LOCKED state made (mask NORM_1|NORM_2|SYNC, tetra_burst_sync.c:117-120)"""
    return events[events["mask"] == 0b1011]


def check_stream_against(records, events, slots, got_records):
    """compare a product result (slots + expanded records) with reference records + search log"""
    ok, msg = records_equal(records, got_records)
    assert ok, msg
    lev = locked_events(events)
    assert lev.size == slots.size, (lev.size, slots.size)
    assert np.array_equal(lev["rc"], slots["find_rc"])
    assert np.array_equal(lev["window"], slots["window"])
    assert np.array_equal(lev["buf_start_bit"], slots["slot_bit"])
    found = lev["rc"] >= 0
    assert np.array_equal(lev["offset"][found], slots["find_off"][found])


# ------------------------------------------------------------------ one stream over several GPUs

def bits_to_symbols(bits, rng, edge_share=0.02):
    """float32 demodulator symbols that float_to_bits (no AFC) slices back to `bits` (even count):
    00 -> (0, 2], 01 -> > 2, 11 -> < -2, 10 -> [-2, 0]; a share of the symbols sits exactly on the
    decision edges 2.0 / 0.0 / -2.0 (float_to_bits.c:33-50: the comparisons are strict)"""
    b = np.ascontiguousarray(bits, dtype=np.uint8).reshape(-1, 2)
    code = b[:, 0] * 2 + b[:, 1]
    u = rng.random(code.size).astype(np.float32)
    lo = np.array([0.0, 2.0, -2.0, -7.0], dtype=np.float32)[code]      # 00, 01, 10, 11
    hi = np.array([2.0, 7.0, 0.0, -2.0], dtype=np.float32)[code]
    f = (lo + (hi - lo) * (0.02 + 0.96 * u)).astype(np.float32)
    edge = rng.random(code.size) < edge_share
    # per code, values on or next to the decision edges and the non-finite ones
    table = np.array([[2.0, 1e-30, 1e-45, 1.9999999],                      # 00: (0, 2]
                      [2.0000002, 7.0, np.inf, 3e38],                      # 01: > 2
                      [-2.0, 0.0, np.nan, -0.0],                           # 10: [-2, 0] and NaN
                      [-2.0000002, -7.0, -np.inf, -3e38]], dtype=np.float32)  # 11: < -2
    pick = rng.integers(0, 4, code.size)
    f[edge] = table[code[edge], pick[edge]]
    return f.astype(np.float32)


REF_FLOAT_TO_BITS = os.path.join(ROOT, "oracle", "_ref", "float_to_bits")


def ref_float_to_bits(sym, tmpdir, args=()):
    """run the reference's own float_to_bits program (compiled unmodified into oracle/_ref) on a file; args: e.g. ("-a",)"""
    fin, fout = os.path.join(tmpdir, "sym.f32"), os.path.join(tmpdir, "sym.bits")
    np.ascontiguousarray(sym, dtype=np.float32).tofile(fin)
    subprocess.check_call([REF_FLOAT_TO_BITS, *args, fin, fout])
    return np.fromfile(fout, dtype=np.uint8)
