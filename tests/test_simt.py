"""Kernel logic and host-side lock state machine, checked on the CPU by compiling the CUDA
sources against the SIMT emulator in tests/simt (test infrastructure, not a product path).
The same scenarios run on the real GPU in test_gpu.py."""
import ctypes as C

import numpy as np
import pytest

import tetra_testlib as T
from tetra_testlib import bits_from_str as B
from test_oracle import SEQS, _stream


def test_time_advance_closed_form(emu, orc):
    lib = emu.lib
    lib.tb200_debug_time_advance.argtypes = [C.POINTER(C.c_uint32)] * 3 + [C.c_uint64]
    rng = np.random.default_rng(0)
    for tn in range(0, 5):
        for fn in (0, 1, 17, 18, 19, 25, 31):
            for mn in (0, 1, 59, 60, 61, 63):
                t = (tn, fn, mn)
                want = {}
                cur = t
                for n in range(0, 6000):
                    if n in (0, 1, 2, 3, 4, 5, 71, 72, 73, 4319, 4320, 4321, 5999) or n % 997 == 0:
                        want[n] = cur
                    cur = orc.time_add_slot(*cur)
                for n, w in want.items():
                    a, b, c = C.c_uint32(tn), C.c_uint32(fn), C.c_uint32(mn)
                    lib.tb200_debug_time_advance(C.byref(a), C.byref(b), C.byref(c), n)
                    assert (a.value, b.value, c.value) == w, (t, n)


def test_find_blind_spot(emu, orc):
    rng = np.random.default_rng(7)
    wins, exp = [], []
    for ts in (T.TS_SYNC, T.TS_NORM_1, T.TS_NORM_2, T.TS_NORM_3):
        sb = B(SEQS[ts])
        for k in range(0, 44):
            for prev in (0, 1):
                for bg in ("zeros", "random"):
                    w = np.zeros(700, np.uint8) if bg == "zeros" else rng.integers(0, 2, 700).astype(np.uint8)
                    w[k:k + sb.size] = sb
                    if k:
                        w[k - 1] = prev
                    wins.append(w)
                    exp.append(orc.find_train_seq(w, 600, 0b1011))
    bits = np.concatenate(wins)
    starts = np.arange(len(wins), dtype=np.uint64) * 700
    rc, off = emu.find_train_seq(bits, starts, np.full(len(wins), 600, np.uint32), 0b1011)
    for i, (r, o) in enumerate(exp):
        assert rc[i] == r and (r < 0 or off[i] == o), (i, (r, o), (rc[i], off[i]))


def test_find_windows(emu, orc):
    """random windows of every length class, including > 1024 (catch-up after a re-lock) and misaligned starts"""
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2, 60000).astype(np.uint8)
    starts, lens = [], []
    for i in range(150):
        st = int(rng.integers(0, 50000)); ln = int(rng.choice([30, 37, 38, 100, 510, 573, 1023, 1024, 1025, 2100, 4096]))
        ts = int(rng.choice([T.TS_SYNC, T.TS_NORM_1, T.TS_NORM_2])); sb = B(SEQS[ts])
        k = int(rng.integers(0, ln + 10))
        if i % 5:
            bits[st + k:st + k + sb.size] = sb
        starts.append(st); lens.append(ln)
    pad = np.concatenate([bits, np.zeros(64, np.uint8)])
    for mask in (0b1011, 0b1000):
        rc, off = emu.find_train_seq(bits, np.array(starts, np.uint64), np.array(lens, np.uint32), mask)
        for i, (st, ln) in enumerate(zip(starts, lens)):
            ln_eff = min(ln, bits.size - st)
            r, o = orc.find_train_seq(pad[st:], ln, mask) if st + ln <= bits.size else (None, None)
            if r is None:
                continue
            assert rc[i] == r and (r < 0 or off[i] == o), (i, st, ln, mask, (r, o), (rc[i], off[i]))


def test_descramble_deinterleave(emu, orc):
    rng = np.random.default_rng(4)
    for K, a in ((120, 11), (216, 101), (432, 103), (168, 13)):
        t5 = rng.integers(0, 2, (9, K)).astype(np.uint8)
        codes = rng.integers(0, 2 ** 32, 9).astype(np.uint32); codes[0] = 3; codes[1] = 0
        got = emu.descramble_deinterleave(t5, codes, K, a)
        for i in range(9):
            want = orc.deinterleave(K, a, orc.scramb_bits(int(codes[i]), t5[i]))
            assert np.array_equal(got[i], want), (K, i)


def _check(emu, orc, bits, chunk=64, **opts):
    orc.reset(); orc.feed(bits, chunk)
    emu.set_options(chunk_bits=chunk, **opts)
    slots, t1, packed = emu.rx_stream_host(bits)
    got = emu.expand_records(slots, t1)
    T.check_stream_against(orc.records(), orc.events(), slots, got)
    # packed output carries the same bits
    unp = ((packed[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(packed.shape[0], 288).astype(np.uint8)
    assert np.array_equal(unp[:, :282], t1[:, :282])
    c = emu.carry()
    assert c.state == orc.rx_state() and c.scramb_init == orc.scramb_init()
    st = emu.stats()             # the counters of tb200_stats against the records
    assert st.slots == slots.size and st.bursts_decoded == int(((slots["flags"] & 3) != 0).sum())
    assert st.blocks == got.size and st.crc_ok_blocks == int(got["crc_ok"][got["lchan"] != T.LC_AACH].sum())
    if slots.size:
        assert (c.tn, c.fn, c.mn) == orc.get_time()
    return slots


@pytest.mark.parametrize("variant", [T.VITERBI_WARP, T.VITERBI_LANE])
def test_stream_mixed(emu, orc, variant):
    bits, _ = _stream(orc, n=120, random_cell=1)
    slots = _check(emu, orc, bits, viterbi=variant, pipeline_slots=0)
    assert slots.size == 119


@pytest.mark.parametrize("variant", [T.VITERBI_WARP, T.VITERBI_LANE])
def test_stream_other_tie_rule(emu, orc, variant):
    """include/tetra_tie_rule.h: with the switch flipped on both sides the kernels still equal the oracle (masked
    four-step form, uniform eight-step form, SB1 pass and warp form all carry the rule)"""
    try:
        orc.set_tie(T.TIE_HIGH_PRED)
        bits, _ = _stream(orc, n=100, random_cell=1, ber_per_65536=2000)
        _check(emu, orc, bits, viterbi=variant, pipeline_slots=0, viterbi_tie=T.TIE_HIGH_PRED)
        bits, _ = _stream(orc, n=140, sb_period=0, ndb2_per_256=0, ber_per_65536=2000, lead_in_bits=5)
        slots = _check(emu, orc, bits, viterbi=variant, pipeline_slots=0, viterbi_tie=T.TIE_HIGH_PRED)
        # and the rule matters on this input: the default setting gives other bits
        emu.set_options(viterbi_tie=T.TIE_LOW_PRED)
        s2, t1b, _ = emu.rx_stream_host(bits)
        emu.set_options(viterbi_tie=T.TIE_HIGH_PRED)
        s1, t1a, _ = emu.rx_stream_host(bits)
        assert not np.array_equal(t1a, t1b)
    finally:
        orc.set_tie(T.TIE_LOW_PRED)
        emu.set_options(viterbi_tie=T.TIE_LOW_PRED)


@pytest.mark.parametrize("form", [0, 1, 2])
def test_stream_lane_forms(orc, form, monkeypatch):
    """the three forms of the lane decode pass (TB200_LANE_FORM: fused kernel, prepare | trellis+finish, prepare | trellis |
    finish) give the oracle's records on a mixed stream with random cells, on uniform SCH/F warps, with the AACH side output
    and with the other tie rule"""
    monkeypatch.setenv("TB200_LANE_FORM", str(form))
    dev = T.B200(emulate=True)
    try:
        bits, _ = _stream(orc, n=150, random_cell=1, ber_per_65536=1500)
        _check(dev, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=0)
        _check(dev, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=33, output=T.OUT_UNPACKED | T.OUT_PACKED)
        bits, _ = _stream(orc, n=140, sb_period=0, ndb2_per_256=0, ber_per_65536=1300, lead_in_bits=5)
        _check(dev, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=0)
        orc.set_tie(T.TIE_HIGH_PRED)
        _check(dev, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=0, viterbi_tie=T.TIE_HIGH_PRED)
    finally:
        orc.set_tie(T.TIE_LOW_PRED)
        dev.close()


def _dev_pieces(dev, buf, n_bits, cuts, unit_bits, host_first=False):
    """feed buf (numpy bytes: the stream in the format options.input names, unit_bits stream bits per byte) through
    tb200_rx_stream_dev in pieces cut at the given bit positions (emulation: host addresses are "device" pointers)"""
    ms = n_bits // 510 + 16
    slots = np.zeros(ms, dtype=T.SLOT_DTYPE); t1 = np.zeros((ms, 288), dtype=np.uint8); pk = np.zeros((ms, 9), dtype=np.uint32)
    got = 0
    edges = [0] + list(cuts) + [n_bits]
    for i in range(len(edges) - 1):
        a, b = edges[i], edges[i + 1]
        flags = (1 if i == 0 else 0) | (2 if i == len(edges) - 2 else 0)
        # byte offset of stream bit a: bytes format 1 B per bit, packed 1 B per 8 bits, symbols 4 B per 2 bits
        boff = {1: a, 8: a // 8, 2: 2 * a}[unit_bits]
        piece = buf.view(np.uint8)[boff:]
        if host_first and i == 0:
            s, t, p = dev.rx_stream_host_raw(piece, b - a, flags)
            n = s.size; slots[got:got + n] = s; t1[got:got + n] = t; pk[got:got + n] = p
        else:
            piece = np.ascontiguousarray(piece)
            n = dev.rx_stream_dev_raw(piece.ctypes.data, b - a, flags, slots[got:].ctypes.data, t1[got:].ctypes.data, pk[got:].ctypes.data, ms - got)
        got += n
    return slots[:got], t1[:got], pk[:got]


@pytest.mark.parametrize("fmt", ["bytes", "packed"])
def test_stream_dev_continuation(emu, orc, fmt):
    """tb200_rx_stream_dev continues a stream across calls (flags = 0) like tb200_rx_stream_host: pieces shorter than the
    receiver's buffer, pieces longer than the staged head, a lock loss, a host call followed by device calls"""
    rng = np.random.default_rng(77)
    bits, _ = _stream(orc, n=260, random_cell=1, ber_per_65536=900)
    bits = bits.copy()
    bits[510 * 120 + 333 + 200:510 * 120 + 333 + 300] ^= 1               # a wiped training sequence: lock loss inside a piece
    n = bits.size
    orc.reset(); orc.feed(bits, 64)
    want, ev = orc.records(), orc.events()
    if fmt == "packed":
        buf = np.packbits(np.concatenate([bits, np.zeros((-n) % 128, np.uint8)]), bitorder="little")
        emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_PACKED)
        q, unit = 128, 8
    else:
        buf = bits
        emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_BYTES)
        q, unit = 1, 1
    try:
        for host_first in (False, True):
            for trial in range(3):
                k = [2, 5, 9][trial]
                cuts = sorted(set(int(c) // q * q for c in rng.integers(1, n, size=k)))
                if trial == 0:
                    cuts = sorted(set(cuts + [40000 // q * q, (40000 + 1024) // q * q, (40000 + 3072) // q * q]))   # short pieces
                cuts = [c for c in cuts if 0 < c < n]
                slots, t1, pk = _dev_pieces(emu, buf, n, cuts, unit, host_first)
                T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
                c = emu.carry()
                assert c.state == orc.rx_state() and c.scramb_init == orc.scramb_init()
        # the stream's end announced by an empty last call (what a caller does that learns of the end afterwards)
        if fmt == "bytes":
            cut = 70001
            ms = n // 510 + 16
            sl = np.zeros(ms, dtype=T.SLOT_DTYPE); t1 = np.zeros((ms, 288), dtype=np.uint8)
            a = np.ascontiguousarray(buf[:cut]); b = np.ascontiguousarray(buf[cut:])
            k0 = emu.rx_stream_dev_raw(a.ctypes.data, cut, 1, sl.ctypes.data, t1.ctypes.data, 0, ms)
            k1 = emu.rx_stream_dev_raw(b.ctypes.data, n - cut, 0, sl[k0:].ctypes.data, t1[k0:].ctypes.data, 0, ms - k0)
            k2 = emu.rx_stream_dev_raw(b.ctypes.data, 0, 2, sl[k0 + k1:].ctypes.data, t1[k0 + k1:].ctypes.data, 0, ms - k0 - k1)
            T.check_stream_against(want, ev, sl[:k0 + k1 + k2], emu.expand_records(sl[:k0 + k1 + k2], t1[:k0 + k1 + k2]))
    finally:
        emu.set_options(input=T.IN_BYTES)


@pytest.mark.parametrize("threads", [1, 3])
def test_stream_host_pack_threads(emu, orc, threads):
    """options.host_pack_threads: the library's host threads pack the caller's bytes before the copy; same records, same
    search log, also piece by piece across calls (the piece that holds the kept tail goes the unpacked way)"""
    bits, _ = _stream(orc, n=180, random_cell=1, ber_per_65536=1200)
    try:
        _check(emu, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=0, host_pack_threads=threads)
        _check(emu, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=37, host_pack_threads=threads)
        _check(emu, orc, bits, chunk=23, viterbi=T.VITERBI_LANE, pipeline_slots=16, host_pack_threads=threads)
        # continuation: three calls
        orc.reset(); orc.feed(bits, 64)
        emu.set_options(chunk_bits=64, pipeline_slots=16, host_pack_threads=threads)
        cuts = [0, 30001, 30001 + 2000, bits.size]
        parts = [emu.rx_stream_host(bits[cuts[i]:cuts[i + 1]], flags=(1 if i == 0 else 0) | (2 if i == 2 else 0)) for i in range(3)]
        slots = np.concatenate([p[0] for p in parts]); t1 = np.concatenate([p[1] for p in parts])
        T.check_stream_against(orc.records(), orc.events(), slots, emu.expand_records(slots, t1))
        # the pool is rebuilt when the thread count changes between calls (new threads must not pick up an old job)
        for th, ps in ((2, 29), (5, 0), (3, 11)):          # (the staging buffers are re-allocated when the piece size grows)
            _check(emu, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=ps, host_pack_threads=th)
    finally:
        emu.set_options(host_pack_threads=0, pipeline_slots=0, chunk_bits=64)


def test_stream_uniform_schf(emu, orc):
    """only SCH/F bursts after the two leading SBs: whole warps of the lane kernel take the unmasked path"""
    bits, _ = _stream(orc, n=200, sb_period=0, ndb2_per_256=0, ber_per_65536=1300, lead_in_bits=5)
    for variant in (T.VITERBI_WARP, T.VITERBI_LANE):
        _check(emu, orc, bits, viterbi=variant, pipeline_slots=0)


@pytest.mark.parametrize("chunk", [1, 7, 64, 100, 296])
def test_stream_chunks(emu, orc, chunk):
    bits, _ = _stream(orc, n=40)
    _check(emu, orc, bits, chunk=chunk, viterbi=T.VITERBI_WARP, pipeline_slots=0)


def test_stream_lock_loss_and_pieces(emu, orc):
    """lock loss inside a pipelined piece: the optimistic pieces are rolled back"""
    bits, _ = _stream(orc, n=150, random_cell=1)
    o = lambda k: 333 + 510 * k
    b2 = bits.copy(); b2[o(60) + 244:o(60) + 266] = 0
    for pieces in (0, 16, 25):
        _check(emu, orc, b2, viterbi=T.VITERBI_WARP, pipeline_slots=pieces)
    assert emu.stats().lock_losses == 1
    b3 = bits.copy(); b3[o(50) + 100:o(50) + 138] = B(SEQS[T.TS_SYNC])
    _check(emu, orc, b3, viterbi=T.VITERBI_WARP, pipeline_slots=16)
    b4 = bits.copy(); b4[o(51) + 30:o(51) + 52] = B(SEQS[T.TS_NORM_2])
    _check(emu, orc, b4, viterbi=T.VITERBI_WARP, pipeline_slots=16)


def test_stream_of_sync_patterns_degrades_like_the_reference(emu, orc, ref):
    """the SYNC sequence overlaps itself at shift 24, so y[0:24] repeated gives ~10 900 SYNC matches per
    2^18 bits - more than the UNLOCKED search's hit list.  The reference cycles lock / unlock and delivers nothing;
    the library must do the same instead of failing the call."""
    y = B(SEQS[T.TS_SYNC])
    bits = np.tile(y[:24], 20000)
    ref.reset(); ref.feed(bits, 64)
    orc.reset(); orc.feed(bits, 64)
    assert ref.records().size == orc.records().size
    assert np.array_equal(ref.events(), orc.events())
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0)
    slots, t1, _ = emu.rx_stream_host(bits)
    got = emu.expand_records(slots, t1)
    T.check_stream_against(ref.records(), ref.events(), slots, got)
    c = emu.carry()
    assert c.state == ref.rx_state()
    assert emu.stats().lock_acquisitions == (ref.events()["mask"] == 0b1000).sum() - (ref.events()[ref.events()["mask"] == 0b1000]["rc"] < 0).sum()


def _feed_runs(rx, lib, bits, runs):
    """the same sequence of read() sizes into a CPU receiver and into the library: runs = [(calls, length), ...]; the
    library gets one call per run of equal-length reads (tb200_rx_stream_host continues the stream, flags = 0)"""
    rx.reset()
    pos, outs = 0, []
    for i, (m, ln) in enumerate(runs):
        part = bits[pos:pos + m * ln]
        pos += part.size
        rx.feed(part, ln)
        lib.set_options(chunk_bits=ln)
        last = i == len(runs) - 1
        flags = (T.TB200_FRESH if i == 0 else 0) | (T.TB200_FINAL if last else 0)
        outs.append(lib.rx_stream_host(part, flags=flags))
    assert pos == bits.size
    slots = np.concatenate([o[0] for o in outs]); t1 = np.concatenate([o[1] for o in outs])
    return slots, t1


def _random_runs(rng, n_bits, max_len=64):
    """reads the way a pipe hands them out: mostly full 64-byte reads, now and then shorter ones of every length"""
    runs, left = [], n_bits
    while left:
        ln = 64 if rng.random() < 0.5 else int(rng.integers(1, max_len + 1))
        m = int(rng.integers(1, 40)) if ln == 64 else int(rng.integers(1, 4))
        m = min(m, left // ln)
        if m == 0:
            ln, m = left, 1          # the last, short read
            if ln > 296:
                ln, m = 64, left // 64
        runs.append((m, ln))
        left -= m * ln
    return runs


def test_stream_variable_read_sizes(emu, ref, orc):
    """tetra-rx on a pipe: read() returns what is there (tetra-rx.c:82-95), so the calls of tetra_burst_sync_in() have
    changing lengths and with them the search windows (bits_in_buf) and the calls that process a slot.  Records, search
    log and final state must equal the reference fed with the same read sizes - also across lock losses."""
    rng = np.random.default_rng(31)
    emu.set_options(viterbi=T.VITERBI_LANE, pipeline_slots=0)
    for case in range(4):
        bits, cfg = _stream(orc, n=90, random_cell=1, sb_period=6, lead_in_bits=int(rng.integers(0, 400)))
        bits = bits.copy()
        if case % 2:
            k = 41
            while orc.gen_kind(cfg, k) == 1:
                k += 1
            o = cfg.lead_in_bits + 510 * k
            bits[o + 244:o + 266] = 0
        runs = _random_runs(rng, bits.size)
        slots, t1 = _feed_runs(ref, emu, bits, runs)
        T.check_stream_against(ref.records(), ref.events(), slots, emu.expand_records(slots, t1))
        c = emu.carry()
        assert c.state == ref.rx_state() and c.scramb_init == ref.scramb_init() and c.calls == sum(m for m, _ in runs)
        if slots.size:
            assert (c.tn, c.fn, c.mn) == ref.get_time()
    emu.set_options(chunk_bits=64)


def _rm_words(orc, rng, n):
    """received words around code words: 0..9 errors, and purely random ones"""
    words = []
    for i in range(n):
        cw = orc.rm3014(int(rng.integers(0, 1 << 14)))
        e = 0
        for b in rng.choice(30, int(rng.integers(0, 10)), replace=False):
            e |= 1 << int(b)
        words.append(cw ^ e if i % 5 else int(rng.integers(0, 1 << 30)))
    return np.array(words, dtype=np.uint32)


def test_rm3014_leaf(emu, orc):
    """tb200_rm3014_decode (syndrome -> coset leader table) equals the exhaustive nearest-code-word search, ties included"""
    rng = np.random.default_rng(8)
    words = _rm_words(orc, rng, 400)
    info, dist, bad = emu.rm3014_decode(words)
    for i, w in enumerate(words):
        wi, wd = orc.rm3014_decode_ml(int(w))
        assert (int(info[i]), int(dist[i])) == (wi, wd), (i, hex(int(w)))
        assert int(bad[i]) == int(wd != 0)


def _check_aach(lib, orc, bits):
    """the AACH side output of the chain: RM-decoded broadcast block of every delivered slot"""
    slots, t1, pk, aach = lib.rx_stream_host_aach(bits)
    kinds = slots["flags"] & 3
    assert (aach[kinds == 0] == 0xffffffff).all()
    n_err = 0
    for i in np.nonzero(kinds)[0]:
        a = int(slots["slot_bit"][i])
        burst = bits[a:a + 510]
        raw = np.concatenate([burst[252:282]]) if kinds[i] == 1 else np.concatenate([burst[230:244], burst[266:282]])
        desc = orc.scramb_bits(int(slots["scrambling_code"][i]), raw)
        word = int("".join(str(int(b)) for b in desc), 2)              # first bit on air = bit 29
        wi, wd = orc.rm3014_decode_ml(word)
        assert (int(aach[i]) & 0x3fff, (int(aach[i]) >> 16) & 0xff, (int(aach[i]) >> 24) & 1) == (wi, wd, int(wd != 0)), i
        # the reference's own AACH type-1 bits are the uncorrected ones: equal when nothing had to be corrected
        off = 60 if kinds[i] == 1 else 0
        uncorrected = int("".join(str(int(b)) for b in t1[i][off:off + 14]), 2)
        assert (uncorrected == wi) or wd > 0
        n_err += int(wd > 0)
    return slots, aach, n_err


def test_stream_aach_rm_decoding(emu, orc):
    bits, cfg = _stream(orc, n=60, random_cell=1, sb_period=3)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0)
    slots, aach, n_err = _check_aach(emu, orc, bits)
    assert n_err <= 2                       # the generator leaves the broadcast block clean (before the first cell is known it is not)
    noisy = bits.copy()
    rng = np.random.default_rng(5)
    for k in range(3, 58):                  # up to three bit errors in every broadcast block: all corrected
        o = 333 + 510 * k
        pos = np.arange(266, 282)            # part of the broadcast block in SYNC and in normal bursts alike
        for p in rng.choice(pos, int(rng.integers(1, 4)), replace=False):
            noisy[o + p] ^= 1
    s2, a2, n_err2 = _check_aach(emu, orc, noisy)
    assert n_err2 > 20
    assert np.array_equal(s2["slot_bit"], slots["slot_bit"])
    ok = (slots["flags"] & 3) != 0
    assert np.array_equal(a2[ok] & 0x3fff, aach[ok] & 0x3fff)       # every damaged block decodes to what the clean one carried
    emu.set_options(viterbi=T.VITERBI_WARP)
    s3, _, _, a3 = emu.rx_stream_host_aach(noisy)
    assert np.array_equal(a2, a3)
    emu.set_options(viterbi=T.VITERBI_LANE)


def test_stream_edge_inputs(emu, orc):
    bits, _ = _stream(orc, n=12)
    rng = np.random.default_rng(3)
    for n in (0, 1, 63, 509, 1019, 1020, 1021, 1531, 333 + 510 * 7 + 123):
        _check(emu, orc, bits[:n], viterbi=T.VITERBI_WARP, pipeline_slots=0)
    _check(emu, orc, rng.integers(0, 2, 9000).astype(np.uint8), viterbi=T.VITERBI_WARP, pipeline_slots=0)
    sb_only, _ = _stream(orc, n=30, sb_period=1, random_cell=1, lead_in_bits=17)
    _check(emu, orc, sb_only, viterbi=T.VITERBI_WARP, pipeline_slots=0)


def test_stream_continuation(emu, orc):
    """feeding the stream in arbitrary pieces gives what one call gives"""
    bits, _ = _stream(orc, n=60, random_cell=1)
    bits[333 + 510 * 20 + 244:333 + 510 * 20 + 266] = 0
    orc.reset(); orc.feed(bits, 64)
    want = orc.records()
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_WARP, pipeline_slots=0)
    rng = np.random.default_rng(5)
    cuts = sorted(set(int(x) for x in rng.integers(1, bits.size, 7))) + [bits.size]
    got = []
    pos = 0
    for i, c in enumerate(cuts):
        flags = (T.TB200_FRESH if i == 0 else 0) | (T.TB200_FINAL if c == bits.size else 0)
        slots, t1, _ = emu.rx_stream_host(bits[pos:c], flags=flags)
        got.append(emu.expand_records(slots, t1))
        pos = c
    ok, msg = T.records_equal(want, np.concatenate(got))
    assert ok, msg


def _uplink_blocks(orc, n, rng, ber):
    """SCH/HU blocks built with the oracle's TX-side functions: 92 type-1 bits + CRC-16 + 4 tail bits -> rate-2/3
    RCPC -> (168, 13) interleaver -> scrambler, then bit errors"""
    t5 = np.zeros((n, 168), np.uint8); codes = rng.integers(1, 2**32, n, dtype=np.uint64).astype(np.uint32)
    for i in range(n):
        t1 = rng.integers(0, 2, 92).astype(np.uint8)
        crc = orc.crc16(t1) ^ 0xffff
        t2 = np.concatenate([t1, [(crc >> (15 - b)) & 1 for b in range(16)], [0, 0, 0, 0]]).astype(np.uint8)
        t3 = orc.punct_2_3(orc.conv_encode(t2), 168)
        t5[i] = orc.scramb_bits(int(codes[i]), orc.interleave(168, 13, t3))
    flips = rng.random(t5.shape) < ber
    return t5 ^ flips.astype(np.uint8), codes


@pytest.mark.parametrize("variant", [T.VITERBI_WARP, T.VITERBI_LANE])
def test_leaf_decode_uplink_and_bbk(emu, orc, variant):
    """the two rows of tetra_blk_param[] the downlink receiver never produces: SCH/HU (168 -> 92) and BBK"""
    rng = np.random.default_rng(21)
    emu.set_options(viterbi=variant)
    t5, codes = _uplink_blocks(orc, 24, rng, 0.02)
    t5[20:] = rng.integers(0, 2, (4, 168))                     # pure noise: ties everywhere
    out, ok = emu.decode_blocks(T.T_SCH_HU, t5, codes)
    for i in range(t5.shape[0]):
        orc.reset(); orc.set_cell(int(codes[i])); orc.tp_sap(T.T_SCH_HU, 0, t5[i])
        r = orc.records()[0]
        assert np.array_equal(r["type1"][:92], out[i]) and int(r["crc_ok"]) == int(ok[i]), i
    assert ok[:20].sum() >= 12
    bb = rng.integers(0, 2, (16, 30)).astype(np.uint8)
    out, ok = emu.decode_blocks(T.T_BBK, bb, codes[:16])
    for i in range(16):
        orc.reset(); orc.set_cell(int(codes[i])); orc.tp_sap(T.T_BBK, 0, bb[i])
        r = orc.records()[0]
        assert np.array_equal(r["type1"][:14], out[i]) and ok[i] == 1
    emu.set_options(viterbi=T.VITERBI_LANE)
