"""The oracle against the reference: known-answer vectors the reference carries, the
reference's own code compiled in place (oracle/_ref), and the committed golden fixtures.
CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import tetra_testlib as T
from tetra_testlib import bits_from_str as B, bits_to_str as S

# crc_test.c:43-57 vector (60 SYNC-PDU bits) and the chain SURVEY.md Appendix B captured
# from the compiled reference functions
KAT_TYPE2 = "00010000101100001011111000000000100000110000011111010011001111011110111100010000"
KAT_TYPE3 = "000001100011111100000101111100110100111110000000111010110110110101001010011010001101110011000010111010001110011100100011"
KAT_TYPE4 = "111001110110101010010101100000111011001100000111010100001111011001101111000101001110111000010011010001011111010010000010"
KAT_TYPE5 = "010110001001111001100100000110010111001101000000111100100101100011001100101101100001111000111100111110101011111000111011"
KAT_MOTHER64 = "0000000000001111101101100101111111111011100100010010001110101111"


def impls(ref, orc):
    return [("ref", ref), ("oracle", orc)]


def test_scrambler_kat(ref, orc):
    for name, x in impls(ref, orc):
        assert S(x.scramb_get_bits(3, 64)) == "1011111111110100111100011001101011000000010001111010001010101110", name
        assert x.scramb_get_init(262, 42, 1) == 0x41802a07, name
        assert x.scramb_get_init(262, 42, 0) == 0x41802a03, name
        assert S(x.scramb_get_bits(0x41802a07, 64)) == "0100011000101001010110010010100000110001000000100001110001011110", name
        assert S(x.scramb_get_bits(0, 16)) == "0" * 16, name


def test_rm3014_kat(ref, orc):
    for name, x in impls(ref, orc):
        assert x.rm3014(0x1001) == 0x10012907, name
        assert x.rm3014(0x3fff) == 0x3fff4abf, name
        assert x.rm3014(0x0001) == 0x000104e7, name
        assert x.rm3014(0x2000) == 0x20009b60, name
    for v in range(0, 1 << 14, 37):
        assert ref.rm3014(v) == orc.rm3014(v)


def test_crc_kat(ref, orc):
    t2 = B(KAT_TYPE2)
    for name, x in impls(ref, orc):
        assert x.crc16(t2[:60]) == 0x210e, name               # crc_test.c:60-62
        assert x.crc16(t2[:76]) == 0x1d0f, name               # TETRA_CRC_OK, tetra_common.h:69
    assert S(t2[60:76]) == "1101111011110001"                  # ~0x210e appended MSB first


def test_chain_kat(ref, orc):
    t2, t3, t4, t5 = B(KAT_TYPE2), B(KAT_TYPE3), B(KAT_TYPE4), B(KAT_TYPE5)
    for name, x in impls(ref, orc):
        m = x.conv_encode(t2)
        assert S(m[:64]) == KAT_MOTHER64, name
        assert np.array_equal(x.punct_2_3(m, 120), t3), name
        assert np.array_equal(x.interleave(120, 11, t3), t4), name
        assert np.array_equal(x.scramb_bits(3, t4), t5), name
        assert np.array_equal(x.deinterleave(120, 11, t4), t3), name
        dp = x.depunct_2_3(t3, 320)
        assert np.array_equal(x.viterbi(dp, 80), t2), name


def test_index_maps(ref, orc):
    # deinterleave source index and depuncture index maps quoted in SURVEY Appendix B
    for name, x in impls(ref, orc):
        for K, a, head in ((120, 11, [11, 22, 33, 44, 55, 66, 77, 88, 99, 110, 1, 12]),
                           (216, 101, [101, 202, 87, 188, 73, 174, 59, 160]),
                           (432, 103, [103, 206, 309, 412, 83, 186, 289, 392])):
            src = np.arange(K, dtype=np.uint8) if K < 256 else None
            # index recovery works on bytes < 256, so probe one-hot instead for K = 432
            idx = []
            for j in range(len(head)):
                hit = [m for m in range(K) if x.deinterleave(K, a, np.eye(1, K, m, dtype=np.uint8)[0])[j]]
                idx.append(hit[0])
            assert idx == head, (name, K)
        one = lambda j: x.depunct_2_3(np.eye(1, 12, j, dtype=np.uint8)[0], 32)
        pos = [int(np.nonzero(one(j) == 1)[0][0]) for j in range(12)]
        assert pos == [0, 1, 4, 8, 9, 12, 16, 17, 20, 24, 25, 28], name


@pytest.mark.parametrize("type2_len,type3_len", [(80, 120), (144, 216), (112, 168), (288, 432)])
def test_punct_roundtrip(ref, orc, type2_len, type3_len):
    """tetra_punct_test tuples that use the 2/3 puncturer (tetra_conv_enc.c:257-267)"""
    mother = (np.arange(type2_len * 4) % 255).astype(np.uint8)
    for name, x in impls(ref, orc):
        t3 = x.punct_2_3(mother, type3_len)
        dp = x.depunct_2_3(t3, type2_len * 4)
        keep = dp != 0xff
        # value 0xff never occurs in `mother`, so every kept symbol must equal the original
        assert keep.sum() == type3_len, name
        assert np.array_equal(dp[keep], mother[keep]), name
    assert np.array_equal(ref.punct_2_3(mother, type3_len), orc.punct_2_3(mother, type3_len))


def _encode(x, t1, bt, code):
    K, N, T1, a = T.BLK[bt]
    t2 = np.zeros(N, np.uint8)
    t2[:T1] = t1
    crc = (~x.crc16(t2[:T1])) & 0xffff
    t2[T1:T1 + 16] = [(crc >> (15 - b)) & 1 for b in range(16)]
    return x.scramb_bits(code, x.interleave(K, a, x.punct_2_3(x.conv_encode(t2), K)))


@pytest.mark.parametrize("bt", [T.T_SB1, T.T_NDB, T.T_SCH_F])
def test_loopback_clean(ref, orc, bt):
    """conv_enc_test.c:336-346: encode -> scramble -> decode must give CRC OK and the payload back"""
    rng = np.random.default_rng(bt)
    K, N, T1, a = T.BLK[bt]
    for _ in range(20):
        t1 = rng.integers(0, 2, T1).astype(np.uint8)
        t5r, t5o = _encode(ref, t1, bt, 3), _encode(orc, t1, bt, 3)
        assert np.array_equal(t5r, t5o)
        for x in (ref, orc):
            t3 = x.deinterleave(K, a, x.scramb_bits(3, t5r))
            t2 = x.viterbi(x.depunct_2_3(t3, 4 * N), N)
            assert x.crc16(t2[:T1 + 16]) == 0x1d0f
            assert np.array_equal(t2[:T1], t1)


@pytest.mark.parametrize("n", [80, 144, 288])
def test_viterbi_noisy_three_restatements(ref, orc, n):
    """oracle Viterbi == acc-style == generic-style restatement of osmo_conv_decode on noisy input
    (tie-breaks matter here; parity with a real libosmocore build is unpinned)"""
    rng = np.random.default_rng(n)
    lib = ref.lib
    K = n * 3 // 2
    for ber in (0.0, 0.01, 0.03, 0.06, 0.15, 0.5):
        for _ in range(40):
            t2 = np.zeros(n, np.uint8)
            t2[:n - 4] = rng.integers(0, 2, n - 4)
            t3 = orc.punct_2_3(orc.conv_encode(t2), K)
            t3 ^= (rng.random(K) < ber).astype(np.uint8)
            mother = orc.depunct_2_3(t3, 4 * n)
            a = orc.viterbi(mother, n)
            b = ref.viterbi(mother, n)                     # reference wrapper + acc-style stand-in
            soft = np.zeros(4 * (n + 4), np.int8)
            soft[:4 * n] = np.where(mother == 0, 127, np.where(mother == 0xff, 0, -127))
            c = np.zeros(n, np.uint8)
            d = np.zeros(n, np.uint8)
            lib.oracle_tetra_cch_decode(0, soft.ctypes.data_as(C.c_void_p), n, c.ctypes.data_as(C.c_void_p))
            lib.oracle_tetra_cch_decode(1, soft.ctypes.data_as(C.c_void_p), n, d.ctypes.data_as(C.c_void_p))
            assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d), (ber,)


def test_viterbi_noisy_kat(ref, orc):
    """SURVEY Appendix B noisy vector (regenerated from the oracle, depends on the tie rule)"""
    t5 = B(KAT_TYPE5)
    for i in (5, 6, 50, 51, 90):
        t5[i] ^= 1
    want = "00010000101100001011111000000000100000110001010100010011001111011110111100010000"
    for name, x in impls(ref, orc):
        t3 = x.deinterleave(120, 11, x.scramb_bits(3, t5))
        t2 = x.viterbi(x.depunct_2_3(t3, 320), 80)
        assert S(t2) == want, name
        assert x.crc16(t2[:76]) == 0xcf22, name


def test_tdma_step(ref, orc):
    for tn in range(0, 6):
        for fn in range(0, 33):
            for mn in range(0, 65):
                assert ref.time_add_slot(tn, fn, mn) == orc.time_add_slot(tn, fn, mn)


SEQS = {T.TS_SYNC: "11000001100111001110100111000001100111", T.TS_NORM_1: "1101000011101001110100",
        T.TS_NORM_2: "0111101001000011011110", T.TS_NORM_3: "1011011100000110101101",
        T.TS_EXT: "100111010000111010011101000011"}


def test_find_blind_spot(ref, orc):
    """tetra_burst.c:288-294 quirk: in[20] never enters the pre-filter"""
    rng = np.random.default_rng(7)
    for ts, seq in SEQS.items():
        sb = B(seq)
        for k in range(0, 48):
            for prev in (0, 1):
                for bg in ("zeros", "random"):
                    w = np.zeros(700, np.uint8) if bg == "zeros" else rng.integers(0, 2, 700).astype(np.uint8)
                    w[k:k + sb.size] = sb
                    if k:
                        w[k - 1] = prev
                    for mask in (0b01011, 0b01000, 0b11111):
                        assert ref.find_train_seq(w, 600, mask) == orc.find_train_seq(w, 600, mask), (ts, k, prev, bg, mask)
    # zeros background: placements 0..20 are missed, 21.. found (the survey's experiment)
    for k in range(0, 40):
        w = np.zeros(700, np.uint8)
        w[k:k + 22] = B(SEQS[T.TS_NORM_1])
        rc, off = orc.find_train_seq(w, 600, 0b01011)
        assert (rc, off) == ((T.TS_NORM_1, k) if k >= 21 else (-1, 0)), k


def test_find_window_edges(ref, orc):
    rng = np.random.default_rng(8)
    for _ in range(300):
        w = rng.integers(0, 2, 800).astype(np.uint8)
        ts = int(rng.choice([T.TS_SYNC, T.TS_NORM_1, T.TS_NORM_2]))
        sb = B(SEQS[ts])
        end = int(rng.integers(30, 700))
        k = int(rng.integers(max(0, end - 60), end + 5))
        w[k:k + sb.size] = sb
        assert ref.find_train_seq(w, end, 0b01011) == orc.find_train_seq(w, end, 0b01011)


def _stream(orc, **kw):
    base = dict(seed=0x7E7A0003, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655,
                random_cell=0, lead_in_bits=333)
    n = kw.pop("n", 200)
    base.update(kw)
    cfg = T.GenCfg(**base)
    return orc.gen_stream(cfg, 0, n), cfg


def _compare_fsm(ref, orc, bits, chunk=64):
    ref.reset(); ref.feed(bits, chunk)
    orc.reset(); orc.feed(bits, chunk)
    ok, msg = T.records_equal(ref.records(), orc.records())
    assert ok, msg
    assert np.array_equal(ref.events(), orc.events())
    assert ref.rx_state() == orc.rx_state()
    assert ref.scramb_init() == orc.scramb_init()
    assert ref.get_time() == orc.get_time()
    return ref.records()


@pytest.mark.parametrize("chunk", [1, 7, 64, 100, 296])
def test_fsm_chunks(ref, orc, chunk):
    bits, _ = _stream(orc, n=60)
    rec = _compare_fsm(ref, orc, bits, chunk)
    assert rec.size > 100


def test_fsm_scenarios(ref, orc):
    bits, cfg = _stream(orc, n=300, random_cell=1)
    rec = _compare_fsm(ref, orc, bits)
    assert rec["crc_ok"].mean() > 0.9
    # lost training sequence -> UNLOCKED -> re-acquire at the next SB
    b2 = bits.copy(); b2[333 + 510 * 100 + 244:333 + 510 * 100 + 266] = 0
    r2 = _compare_fsm(ref, orc, b2)
    assert r2.size < rec.size
    # SYNC sequence at the wrong offset inside a normal burst -> UNLOCKED (tetra_burst_sync.c:123-128)
    b3 = bits.copy(); b3[333 + 510 * 50 + 100:333 + 510 * 50 + 138] = B(SEQS[T.TS_SYNC])
    _compare_fsm(ref, orc, b3)
    # early NORM hit -> burst dropped, still LOCKED (tetra_burst_sync.c:133-137)
    b4 = bits.copy(); b4[333 + 510 * 51 + 30:333 + 510 * 51 + 52] = B(SEQS[T.TS_NORM_2])
    r4 = _compare_fsm(ref, orc, b4)
    assert 0 < rec.size - r4.size <= 3
    # all-SB stream with per-burst random cells (config 4 shape)
    b5, _ = _stream(orc, n=120, sb_period=1, random_cell=1)
    _compare_fsm(ref, orc, b5)
    # garbage
    rng = np.random.default_rng(3)
    _compare_fsm(ref, orc, rng.integers(0, 2, 20000).astype(np.uint8))
    # short / ragged / empty
    for n in (0, 1, 63, 509, 1019, 1020, 1021, 1531):
        _compare_fsm(ref, orc, bits[:n])
    _compare_fsm(ref, orc, bits[:333 + 510 * 7 + 123])


def test_bnch_lchan(ref, orc):
    """SB2 is BNCH when fn == 18 and tn == 4 - ((mn+3) % 4) (tetra_lower_mac.c:122-127)"""
    # burst k announces tn = k%4+1, fn = (k/4)%18+1, mn = (k/72)%60+1: k = 68..71 have fn 18
    bits, _ = _stream(orc, n=80, sb_period=1, lead_in_bits=0, ber_per_65536=0)
    rec = _compare_fsm(ref, orc, bits)
    assert (rec["lchan"] == T.LC_BNCH).sum() >= 1


def test_generator_matches_reference_tx(ref, orc):
    """the oracle's stream generator against the reference's own TX functions"""
    cfg = T.GenCfg(seed=11, sb_period=3, lead_sb=2, ndb2_per_256=100, ber_per_65536=0, random_cell=1, lead_in_bits=0)
    bits = orc.gen_stream(cfg, 0, 12)
    orc.reset(); orc.feed(np.concatenate([bits, bits[:510 * 2]]))
    rec = orc.records()
    assert rec["crc_ok"].all()      # burst 0 gives lock, burst 1 (an SB) sets the cell code
    # rebuild every burst from the decoded payloads with the reference TX chain
    by_slot = {}
    for r in rec:
        by_slot.setdefault(int(r["slot_bit"]), []).append(r)
    checked = 0
    for slot_bit, rs in by_slot.items():
        k = slot_bit // 510
        if k >= 12:
            continue
        burst = bits[510 * k:510 * (k + 1)]
        kind = orc.gen_kind(cfg, k)
        code = int(rs[1]["scrambling_code"])
        bb = np.array([(ref.rm3014(int(S(rs[0 if kind != T.TS_SYNC else 1]["type1"][:14]), 2)) >> (29 - i)) & 1 for i in range(30)], np.uint8)
        bb = ref.scramb_bits(code, bb)
        if kind == T.TS_SYNC:
            sb = _encode(ref, rs[0]["type1"][:60], T.T_SB1, 3)
            bkn = _encode(ref, rs[2]["type1"][:124], T.T_SB2, code)
            want = ref.build_sync_burst(sb, bb, bkn)
        elif kind == T.TS_NORM_1:
            f = _encode(ref, rs[1]["type1"][:268], T.T_SCH_F, code)
            want = ref.build_norm_burst(f[:216], bb, f[216:], 0)
        else:
            b1 = _encode(ref, rs[1]["type1"][:124], T.T_NDB, code)
            b2 = _encode(ref, rs[2]["type1"][:124], T.T_NDB, code)
            want = ref.build_norm_burst(b1, bb, b2, 1)
        # phase adjustment bits (12,13 and 498,499) come from an out-of-bounds table read in the
        # reference builder (tetra_burst.c:163, index without PHASE()); the generator leaves them 0
        m = np.ones(510, bool); m[[12, 13, 498, 499]] = False
        assert np.array_equal(burst[m], want[m]), k
        checked += 1
    assert checked >= 10


def test_uplink_and_bbk_rows(ref, orc):
    """tetra_blk_param[] rows the downlink slicer never produces (SCH/HU) or that skip the channel decoder (BBK):
    the restatement against the reference's tp_sap_udata_ind called directly"""
    from test_simt import _uplink_blocks
    rng = np.random.default_rng(22)
    t5, codes = _uplink_blocks(orc, 30, rng, 0.03)
    t5[24:] = rng.integers(0, 2, (6, 168))
    for i in range(t5.shape[0]):
        # the reference learns its cell code only from an SB1: descramble here and feed both with code 0 instead
        plain = orc.scramb_bits(int(codes[i]), t5[i])
        orc.reset(); ref.reset()
        ref.tp_sap(T.T_SCH_HU, 0, plain); orc.tp_sap(T.T_SCH_HU, 0, plain)
        ok, msg = T.records_equal(ref.records(), orc.records())
        assert ok, (i, msg)
    for i in range(8):
        bb = rng.integers(0, 2, 30).astype(np.uint8)
        orc.reset(); ref.reset()
        ref.tp_sap(T.T_BBK, 0, bb); orc.tp_sap(T.T_BBK, 0, bb)
        ok, msg = T.records_equal(ref.records(), orc.records())
        assert ok, (i, msg)


@pytest.mark.parametrize("tie", [T.TIE_LOW_PRED, T.TIE_HIGH_PRED])
def test_viterbi_tie_switch(ref, orc, tie):
    """include/tetra_tie_rule.h: the one switch moves the oracle port and both restatements of osmo_conv_decode
    together (the reference wrapper runs on the acc-style stand-in); the two settings really differ on noisy blocks"""
    rng = np.random.default_rng(99 + tie)
    lib = ref.lib
    differ = 0
    try:
        for n in (80, 144, 288):
            K = n * 3 // 2
            for _ in range(60):
                t2 = np.zeros(n, np.uint8)
                t2[:n - 4] = rng.integers(0, 2, n - 4)
                t3 = orc.punct_2_3(orc.conv_encode(t2), K)
                t3 ^= (rng.random(K) < 0.03).astype(np.uint8)
                mother = orc.depunct_2_3(t3, 4 * n)
                orc.set_tie(tie); ref.set_tie(tie)
                a = orc.viterbi(mother, n)
                b = ref.viterbi(mother, n)
                soft = np.zeros(4 * (n + 4), np.int8)
                soft[:4 * n] = np.where(mother == 0, 127, np.where(mother == 0xff, 0, -127))
                c = np.zeros(n, np.uint8)
                d = np.zeros(n, np.uint8)
                lib.oracle_tetra_cch_decode(0, soft.ctypes.data_as(C.c_void_p), n, c.ctypes.data_as(C.c_void_p))
                lib.oracle_tetra_cch_decode(1, soft.ctypes.data_as(C.c_void_p), n, d.ctypes.data_as(C.c_void_p))
                assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d), (tie, n)
                orc.set_tie(1 - tie)
                differ += int(not np.array_equal(a, orc.viterbi(mother, n)))
    finally:
        orc.set_tie(T.TIE_LOW_PRED); ref.set_tie(T.TIE_LOW_PRED)
    assert differ > 20          # the rule changes 70 % of the blocks at 3 % BER


def test_pin_conv_program_selfcheck(tmp_path):
    """oracle/pin_conv.c (`make -C oracle pin-libosmocore`) compiles and, linked against the stand-in in the role of
    the library, reports the compiled-in tie rule as matching and the other rule as differing"""
    import subprocess
    exe = str(tmp_path / "pin_self")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(T.ORACLE_DIR, "stubs"), "-o", exe,
                           os.path.join(T.ORACLE_DIR, "pin_conv.c"), os.path.join(T.ORACLE_DIR, "osmo_standin.c")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "PINNED" in r.stdout and " on 0 (ties keep s>>1)" in r.stdout and "/ 0 (ties" not in r.stdout, r.stdout
    r = subprocess.run(["make", "-s", "-C", T.ORACLE_DIR, "pin-libosmocore"], capture_output=True, text=True)
    assert r.returncode == 0 and ("UNPINNED" in r.stdout or "PINNED" in r.stdout), r.stdout + r.stderr


def test_rm3014_ml_decoder(ref, orc):
    """the brute-force RM(30,14) decoder of the oracle: the code has minimum distance 8 (checked here over all 2^14 code
    words of the REFERENCE's tetra_rm3014_compute), so every pattern of up to three errors decodes back"""
    cws = np.array([ref.rm3014(i) for i in range(1 << 14)], dtype=np.uint32)
    assert np.array_equal(cws, np.array([orc.rm3014(i) for i in range(1 << 14)], dtype=np.uint32))
    w = np.array([bin(int(x)).count("1") for x in cws[1:]])
    assert w.min() == 8                                            # linear code: minimum distance = lightest non-zero word
    rng = np.random.default_rng(3014)
    for _ in range(300):
        info = int(rng.integers(0, 1 << 14))
        ne = int(rng.integers(0, 4))
        e = 0
        for b in rng.choice(30, ne, replace=False):
            e |= 1 << int(b)
        got, d = orc.rm3014_decode_ml(int(cws[info]) ^ e)
        assert (got, d) == (info, ne)
