"""The drop-in shim (osmo-tetra_b200/host/tetra_shim.c) compiled against the reference's own headers and
linked to the SIMT-emulation build: tetra_burst_sync_in() in 64-byte reads must produce, at
upper_mac_prim_recv(), the records the reference's PHY + lower MAC produce.  Needs /root/reference (headers)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import tetra_testlib as T
from test_oracle import _stream

REF_SRC = "/root/reference/src"
BUILD = os.path.join(T.ROOT, "tests", "simt", "_build")

RECORDER = r'''
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <osmocom/core/msgb.h>
#include <tetra_common.h>
#include <tetra_prim.h>
#include "oracle_records.h"
static struct tb_record *g_rec; static size_t g_n, g_cap;
int upper_mac_prim_recv(struct osmo_prim_hdr *op, void *priv)
{
	struct tetra_tmvsap_prim *t = (struct tetra_tmvsap_prim *)op;
	struct tmv_unitdata_param *u = &t->u.unitdata;
	if (g_n == g_cap) { g_cap = g_cap ? 2 * g_cap : 1024; g_rec = realloc(g_rec, g_cap * sizeof(*g_rec)); }
	struct tb_record *r = &g_rec[g_n++];
	memset(r, 0, sizeof(*r));
	r->lchan = u->lchan; r->crc_ok = u->crc_ok; r->blk_num = u->blk_num; r->scrambling_code = u->scrambling_code;
	r->tn = u->tdma_time.tn; r->fn = u->tdma_time.fn; r->mn = u->tdma_time.mn;
	r->type1_len = msgb_l1len(op->msg);
	memcpy(r->type1, op->msg->l1h, r->type1_len);
	return -1;
}
size_t shimtest_n(void) { return g_n; }
const struct tb_record *shimtest_rec(void) { return g_rec; }
'''


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference headers not present")
def test_shim_delivers_reference_primitives(orc):
    simt = T.build_simt()
    os.makedirs(BUILD, exist_ok=True)
    rec_c = os.path.join(BUILD, "shim_recorder.c")
    open(rec_c, "w").write(RECORDER)
    so = os.path.join(BUILD, "libshimtest.so")
    subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-shared", "-I" + REF_SRC, "-I" + os.path.join(T.ROOT, "oracle", "stubs"),
                           "-I" + os.path.join(T.ROOT, "oracle"), "-I" + os.path.join(T.ROOT, "include"),
                           os.path.join(T.ROOT, "osmo-tetra_b200", "host", "tetra_shim.c"), rec_c, simt,
                           "-Wl,-rpath," + os.path.dirname(simt), "-o", so])
    bits, _ = _stream(orc, n=90, random_cell=1)
    bits[333 + 510 * 40 + 244:333 + 510 * 40 + 266] = 0
    orc.reset(); orc.feed(bits, 64)
    want = orc.records()
    os.environ["TETRA_B200_BATCH_BITS"] = "16384"          # several GPU batches
    lib = C.CDLL(so)
    class Trs(C.Structure):
        _fields_ = [("state", C.c_int), ("bits_in_buf", C.c_uint), ("bitbuf", C.c_uint8 * 4096),
                    ("start", C.c_uint), ("next", C.c_uint), ("priv", C.c_void_p)]
    trs = Trs()
    for pos in range(0, bits.size, 64):
        chunk = np.ascontiguousarray(bits[pos:pos + 64])
        lib.tetra_burst_sync_in(C.byref(trs), chunk.ctypes.data_as(C.c_void_p), chunk.size)
    lib.tetra_b200_shim_flush()
    lib.shimtest_n.restype = C.c_size_t
    lib.shimtest_rec.restype = C.c_void_p
    n = lib.shimtest_n()
    got = np.frombuffer(C.string_at(lib.shimtest_rec(), n * 288), dtype=T.RECORD_DTYPE).copy()
    got["slot_bit"] = want["slot_bit"][:n] if n == want.size else 0      # the primitive does not carry the bit position
    ok, msg = T.records_equal(want, got)
    assert ok, msg
    assert trs.state == orc.rx_state()


@pytest.mark.skipif(not (os.path.isdir(REF_SRC) and T.have_ref()), reason="reference build not present")
def test_shim_l0_only_feeds_reference_lower_mac(orc):
    """Drop-in depth A: only the PHY on the 'GPU' (SIMT-emulation build).  The shim hands every
    delivered slot to the REFERENCE's own tetra_burst_rx_cb (phy/tetra_burst.o) and from there into the
    reference's tp_sap_udata_ind / lower MAC (compiled in place, oracle/_ref/obj); the records at
    upper_mac_prim_recv must be those of the all-reference receiver, lock loss and dropped bursts included."""
    import glob
    simt = T.build_simt()
    os.makedirs(BUILD, exist_ok=True)
    objdir = os.path.join(T.ROOT, "oracle", "_ref", "obj")
    objs = [o for o in glob.glob(os.path.join(objdir, "**", "*.o"), recursive=True) if not o.endswith("tetra_burst_sync.o")]
    so = os.path.join(BUILD, "libshim_l0.so")
    subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-shared", "-DTETRA_B200_SHIM_L0_ONLY", "-I" + REF_SRC,
                           "-I" + os.path.join(T.ROOT, "oracle", "stubs"), "-I" + os.path.join(T.ROOT, "include"),
                           os.path.join(T.ROOT, "osmo-tetra_b200", "host", "tetra_shim.c")] + objs +
                          [simt, "-Wl,--wrap=tetra_find_train_seq", "-Wl,-rpath," + os.path.dirname(simt), "-o", so])
    bits, _ = _stream(orc, n=120, random_cell=1, sb_period=7)
    bits[333 + 510 * 40 + 244:333 + 510 * 40 + 266] = 0           # a wiped training sequence: lock loss + re-acquisition
    full = T.Ref()
    full.reset(); full.feed(bits, 64)
    want = full.records().copy()
    os.environ["TETRA_B200_BATCH_BITS"] = "12288"          # several GPU batches, slots straddling them
    lib = C.CDLL(so)
    lib.ref_feed.restype = C.c_long
    lib.ref_feed.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_int]
    lib.ref_rm3014_init()
    lib.ref_reset()
    b = np.ascontiguousarray(bits)
    lib.ref_feed(b.ctypes.data_as(C.c_void_p), b.size, 64, 1)
    lib.tetra_b200_shim_flush()
    lib.ref_num_records.restype = C.c_size_t
    lib.ref_records.restype = C.c_void_p
    n = lib.ref_num_records()
    got = np.frombuffer(C.string_at(lib.ref_records(), n * 288), dtype=T.RECORD_DTYPE).copy()
    assert n == want.size
    got["slot_bit"] = want["slot_bit"]                     # bitbuf_start_bitnum is not visible at this depth
    ok, msg = T.records_equal(want, got)
    assert ok, msg
    assert int(want["crc_ok"].sum()) > 100


TRAFFIC_RECORDER = RECORDER.replace('''	return -1;
}''', '''	if (priv && u->lchan == TETRA_LC_AACH) {
		struct tetra_mac_state *tms = priv;
		TB_EMULATE_RX_AACH(tms->cur_burst, op->msg->l1h, u->tdma_time.fn);
	}
	return -1;
}
static struct tetra_mac_state g_tms;
void *shimtest_tms(void) { return &g_tms; }
void shimtest_set_dumpdir(const char *d) { g_tms.dumpdir = strdup(d); }''', 1)


def _plant_aach(ref, bits, lead, k, two_block, info14, code):
    """overwrite the broadcast block of normal burst k with the RM(30,14) code word of info14, scrambled with the cell code"""
    word = ref.rm3014(info14)
    cw = np.array([(word >> (29 - i)) & 1 for i in range(30)], dtype=np.uint8)
    cw = ref.scramb_bits(code, cw)
    a = lead + 510 * k
    bits[a + 230:a + 244] = cw[:14]
    bits[a + 266:a + 282] = cw[14:]


@pytest.mark.parametrize("ndb2", [0, 96])
@pytest.mark.skipif(not (os.path.isdir(REF_SRC) and T.have_ref()), reason="reference build not present")
def test_shim_follows_the_traffic_feedback(orc, tmp_path, ndb2):
    """tms->cur_burst.is_traffic, set by the upper MAC when an AACH marks a slot as traffic
    (tetra_upper_mac.c:444-452), makes the reference lower MAC divert SCH/F and un-stolen second blocks to
    its dump files instead of delivering them (tetra_lower_mac.c:190-241).  The shim must withhold exactly
    the same primitives.  Both sides run with a recorder that emulates that feedback (oracle_records.h)."""
    ref = T.Ref()
    cfg = T.GenCfg(seed=0x7E7A0077, sb_period=9, lead_sb=2, ndb2_per_256=ndb2, ber_per_65536=300, random_cell=0, lead_in_bits=123)
    n = 120
    bits = orc.gen_stream(cfg, 0, n).copy()
    code = ref.scramb_get_init(262, 42, 1)
    rng = np.random.default_rng(9)
    marked = 0
    for k in range(3, n - 1):
        if orc.lib.orc_gen_kind(C.byref(cfg), k) == T.TS_SYNC or rng.random() > 0.4:
            continue
        hdr = int(rng.integers(0, 4))
        usage = int(rng.choice([0, 2, 3, 4, 9, 40]))
        info = (hdr << 12) | (usage << 6) | int(rng.integers(0, 64))
        _plant_aach(ref, bits, 123, k, None, info, code)
        marked += hdr != 0 and usage > 3
    assert marked > 5
    # the all-reference receiver with the emulated feedback and a dump directory
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    ref.reset()
    ref.lib.ref_set_feedback(1, str(ref_dir).encode())
    ref.feed(bits, 64)
    want = ref.records().copy()
    ref.lib.ref_set_feedback(0, None)
    plain = T.Ref(); plain.reset(); plain.feed(bits, 64)
    assert 0 < want.size < plain.records().size               # the feedback really withheld primitives
    assert any(f.startswith("traffic_") for f in os.listdir(ref_dir))
    # the shim (PHY + lower MAC on the emulated GPU) with the same feedback in its recorder
    simt = T.build_simt()
    os.makedirs(BUILD, exist_ok=True)
    rec_c = os.path.join(BUILD, "shim_recorder_traffic.c")
    open(rec_c, "w").write(TRAFFIC_RECORDER)
    so = os.path.join(BUILD, f"libshimtest_traffic{ndb2}.so")        # a fresh library per case: the shim keeps one receiver per process
    subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-shared", "-I" + REF_SRC, "-I" + os.path.join(T.ROOT, "oracle", "stubs"),
                           "-I" + os.path.join(T.ROOT, "oracle"), "-I" + os.path.join(T.ROOT, "include"),
                           os.path.join(T.ROOT, "osmo-tetra_b200", "host", "tetra_shim.c"), rec_c, simt,
                           "-Wl,-rpath," + os.path.dirname(simt), "-o", so])
    os.environ["TETRA_B200_BATCH_BITS"] = "16384"
    lib = C.CDLL(so)
    lib.shimtest_tms.restype = C.c_void_p
    class Trs(C.Structure):
        _fields_ = [("state", C.c_int), ("bits_in_buf", C.c_uint), ("bitbuf", C.c_uint8 * 4096),
                    ("start", C.c_uint), ("next", C.c_uint), ("priv", C.c_void_p)]
    trs = Trs()
    trs.priv = lib.shimtest_tms()
    shim_dir = tmp_path / "shim"
    shim_dir.mkdir()
    lib.shimtest_set_dumpdir(str(shim_dir).encode())
    for pos in range(0, bits.size, 64):
        chunk = np.ascontiguousarray(bits[pos:pos + 64])
        lib.tetra_burst_sync_in(C.byref(trs), chunk.ctypes.data_as(C.c_void_p), chunk.size)
    lib.tetra_b200_shim_flush()
    lib.shimtest_n.restype = C.c_size_t
    lib.shimtest_rec.restype = C.c_void_p
    m = lib.shimtest_n()
    got = np.frombuffer(C.string_at(lib.shimtest_rec(), m * 288), dtype=T.RECORD_DTYPE).copy()
    assert m == want.size
    got["slot_bit"] = want["slot_bit"]
    ok, msg = T.records_equal(want, got)
    assert ok, msg
    # the traffic dump of the SCH/F-shaped slots: same files, same bytes (this stream has no two-block slots,
    # whose dump the reference fills from uninitialised memory)
    if ndb2:
        return
    ref_files = sorted(f for f in os.listdir(ref_dir) if f.startswith("traffic_"))
    assert ref_files and ref_files == sorted(os.listdir(shim_dir))
    for f in ref_files:
        assert open(os.path.join(ref_dir, f), "rb").read() == open(os.path.join(shim_dir, f), "rb").read(), f


@pytest.mark.skipif(not (os.path.isdir(REF_SRC) and T.have_ref()), reason="reference build not present")
@pytest.mark.parametrize("batch", ["0", "9000"])
def test_shim_variable_reads_and_return_values(orc, ref, batch):
    """tetra-rx on a pipe (tetra-rx.c:82-95, src/receiver1udp): read() hands out what is there, so the lengths of the
    tetra_burst_sync_in() calls change.  The shim must deliver the reference's primitives for the same call lengths, and
    with TETRA_B200_BATCH_BITS=0 (decode on every call) return the reference's own value from every call
    (tetra_burst_sync.c:77-78,93-94: len / 0 / -1)."""
    from test_simt import _random_runs
    simt = T.build_simt()
    os.makedirs(BUILD, exist_ok=True)
    rec_c = os.path.join(BUILD, "shim_recorder.c")
    open(rec_c, "w").write(RECORDER)
    so = os.path.join(BUILD, f"libshimtest_var{batch}.so")
    subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-shared", "-I" + REF_SRC, "-I" + os.path.join(T.ROOT, "oracle", "stubs"),
                           "-I" + os.path.join(T.ROOT, "oracle"), "-I" + os.path.join(T.ROOT, "include"),
                           os.path.join(T.ROOT, "osmo-tetra_b200", "host", "tetra_shim.c"), rec_c, simt,
                           "-Wl,-rpath," + os.path.dirname(simt), "-o", so])
    rng = np.random.default_rng(12)
    bits, cfg = _stream(orc, n=70, random_cell=1, sb_period=6, lead_in_bits=1500)
    bits = bits.copy()
    bits[1500 + 510 * 33 + 200:1500 + 510 * 33 + 300] = 0          # lock is lost once: UNLOCKED searches that fail return -1
    lens = np.array([ln for m, ln in _random_runs(rng, bits.size) for _ in range(m)], dtype=np.uint32)
    assert lens.sum() == bits.size
    want_rc = np.zeros(lens.size, dtype=np.int32)
    ref.reset()
    ref.lib.ref_feed_calls.restype = C.c_long
    ref.lib.ref_feed_calls.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    b = np.ascontiguousarray(bits)
    ref.lib.ref_feed_calls(b.ctypes.data, lens.ctypes.data, lens.size, want_rc.ctypes.data, 1)
    want = ref.records()
    assert set(np.unique(want_rc)) >= {-1, 0}                       # the stream really exercises all three values
    os.environ["TETRA_B200_BATCH_BITS"] = batch
    os.environ["TETRA_B200_FLUSH_MS"] = "0"
    lib = C.CDLL(so)

    class Trs(C.Structure):
        _fields_ = [("state", C.c_int), ("bits_in_buf", C.c_uint), ("bitbuf", C.c_uint8 * 4096),
                    ("start", C.c_uint), ("next", C.c_uint), ("priv", C.c_void_p)]
    trs = Trs()
    got_rc = np.zeros(lens.size, dtype=np.int32)
    pos = 0
    for i, ln in enumerate(lens):
        chunk = np.ascontiguousarray(bits[pos:pos + ln])
        pos += int(ln)
        got_rc[i] = lib.tetra_burst_sync_in(C.byref(trs), chunk.ctypes.data_as(C.c_void_p), int(ln))
    lib.tetra_b200_shim_flush()
    lib.shimtest_n.restype = C.c_size_t
    lib.shimtest_rec.restype = C.c_void_p
    n = lib.shimtest_n()
    got = np.frombuffer(C.string_at(lib.shimtest_rec(), n * 288), dtype=T.RECORD_DTYPE).copy()
    assert n == want.size
    got["slot_bit"] = want["slot_bit"]
    ok, msg = T.records_equal(want, got)
    assert ok, msg
    assert trs.state == ref.rx_state()
    if batch == "0":
        assert np.array_equal(got_rc, want_rc)
    else:
        assert np.array_equal(got_rc, lens.astype(np.int32))
    del os.environ["TETRA_B200_FLUSH_MS"]
