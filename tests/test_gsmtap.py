"""GSMTAP framing of the decoded blocks.  The reference builds one frame per CRC-good
primitive in tetra_gsmtap_makemsg (tetra_gsmtap.c:31-63, called from tetra_upper_mac.c:480-488); the oracle
restates it, oracle/_ref runs the reference's own tetra_gsmtap.c (its UDP socket replaced by a capture
buffer), and tb200_gsmtap_pack does it on the device from the slot records + packed type-1 words.
CPU: oracle against the reference and a hand-built frame; the CUDA kernels under the SIMT emulator.
GPU (-m gpu): the same through the real library, host and device pointers, plus a 10^6-slot size check."""
import ctypes as C

import numpy as np
import pytest

import tetra_testlib as T
from test_fuzz import make_case
from test_oracle import _stream


def _records(orc, bits, chunk=64):
    orc.reset(); orc.feed(bits, chunk)
    return orc.records()


def test_oracle_frame_by_hand(orc):
    """one SCH/F record, every header byte spelled out"""
    rec = np.zeros(2, dtype=T.RECORD_DTYPE)
    rng = np.random.default_rng(1)
    t1 = rng.integers(0, 2, 268, dtype=np.uint8)
    rec[0]["lchan"] = 1; rec[0]["crc_ok"] = 1; rec[0]["tn"] = 3; rec[0]["fn"] = 17; rec[0]["mn"] = 59
    rec[0]["type1_len"] = 268; rec[0]["type1"][:268] = t1
    rec[1] = rec[0]; rec[1]["crc_ok"] = 0                       # a wrong CRC produces no frame
    fr, n = orc.gsmtap_frames(rec)
    assert n == 1 and fr.size == 16 + 34
    fnum = 59 * 18 + 17
    assert list(fr[:16]) == [2, 4, 5, 2, 0, 0, 0, 0, 0, 0, fnum >> 8, fnum & 0xff, 5, 0, 0, 0]
    want = np.packbits(np.concatenate([t1, np.zeros(4, dtype=np.uint8)]))   # MSB first, zero padded
    assert np.array_equal(fr[16:], want)


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_oracle_matches_reference_gsmtap(orc, ref, seed):
    """the restatement against the reference's own tetra_gsmtap.c on damaged streams"""
    bits, chunk = make_case(orc, seed, 120)
    rec = _records(orc, bits, chunk)
    assert rec.size > 100
    a, na = orc.gsmtap_frames(rec)
    b, nb = ref.gsmtap_frames(rec)
    assert na == nb == int(rec["crc_ok"].sum())
    assert np.array_equal(a, b)


def _check(g, orc, bits):
    want_rec = _records(orc, bits)
    slots, t1, packed = g.rx_stream_host(bits)
    rec = g.expand_records(slots, t1)
    assert rec.size == want_rec.size
    want, nf = orc.gsmtap_frames(want_rec)
    got, off, n = g.gsmtap_pack(slots, packed)
    assert n == nf
    assert np.array_equal(got, want)
    # slot offsets: frames of slot i = the frames of its CRC-good records
    per_rec = np.where(rec["crc_ok"] != 0, 16 + (rec["type1_len"].astype(np.int64) + 7) // 8, 0)
    per_slot = np.zeros(slots.size, dtype=np.int64)
    kinds = slots["flags"] & 3
    nblk = np.where(kinds == 0, 0, np.where(kinds == 2, 2, 3))
    first = np.concatenate([[0], np.cumsum(nblk)])
    csum = np.concatenate([[0], np.cumsum(per_rec)])
    per_slot = csum[first[1:]] - csum[first[:-1]]
    assert np.array_equal(off.astype(np.int64), np.concatenate([[0], np.cumsum(per_slot)]))
    return slots, packed, want


@pytest.mark.parametrize("n,kw", [(40, dict(sb_period=4, ndb2_per_256=128, ber_per_65536=1500, random_cell=1)),
                                  (700, dict(sb_period=7, ndb2_per_256=64, ber_per_65536=2500, random_cell=1, lead_in_bits=333)),
                                  (257, dict(sb_period=0, ndb2_per_256=0, ber_per_65536=0))])
def test_gsmtap_pack_emulated(emu, orc, n, kw):
    """mixed SB / SCH-F / two-half-slot streams with wrong CRCs: byte-identical frames, tile edges (256 slots) included"""
    bits, _ = _stream(orc, n=n, **kw)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_UNPACKED | T.OUT_PACKED)
    _check(emu, orc, bits)


def test_gsmtap_pack_emulated_fuzz(emu, orc):
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_UNPACKED | T.OUT_PACKED)
    for seed in (11, 12):
        bits, _ = make_case(orc, seed, 150)
        _check(emu, orc, bits)


def test_gsmtap_pack_edges(emu):
    """no slots, a buffer that is too small, the size query"""
    empty = np.zeros(0, dtype=T.SLOT_DTYPE)
    fr, off, n = emu.gsmtap_pack(empty, np.zeros((0, 9), dtype=np.uint32))
    assert fr.size == 0 and n == 0
    slots = np.zeros(3, dtype=T.SLOT_DTYPE)
    slots["flags"] = [2 | 4, 0, 3 | 8]                          # SCH/F good, dropped slot, two halves with only BLK2 good
    slots["time"] = 1 | 5 << 3 | 7 << 8
    packed = np.zeros((3, 9), dtype=np.uint32)
    nf = C.c_uint64(0)
    need = emu.lib.tb200_gsmtap_pack(emu.h, T._ptr(slots), None, 3, None, 0, None, C.byref(nf), 0)
    assert need == (18 + 50) + 0 + (18 + 32) and nf.value == 4
    buf = np.zeros(need, dtype=np.uint8)
    rc = emu.lib.tb200_gsmtap_pack(emu.h, T._ptr(slots), T._ptr(packed), 3, T._ptr(buf), need - 2, None, None, 0)
    assert rc < 0 and "bytes" in emu.err()
    rc = emu.lib.tb200_gsmtap_pack(emu.h, T._ptr(slots), T._ptr(packed), 3, T._ptr(buf), need, None, None, 0)
    assert rc == need
    assert list(buf[:4]) == [2, 4, 5, 0] and buf[12] == 2 and buf[18 + 12] == 5 and buf[68 + 12] == 2 and buf[68 + 18 + 12] == 0


# ------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("n,kw", [(3000, dict(sb_period=5, ndb2_per_256=96, ber_per_65536=2000, random_cell=1, lead_in_bits=333)),
                                  (513, dict(sb_period=0, ndb2_per_256=0, ber_per_65536=655))])
def test_gsmtap_pack_gpu(gpu, orc, n, kw):
    import torch
    bits, _ = _stream(orc, n=n, **kw)
    gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_UNPACKED | T.OUT_PACKED)
    slots, packed, want = _check(gpu, orc, bits)
    # device pointers, output at an odd 2-byte phase
    d_slots = torch.from_numpy(slots.view(np.uint8).copy()).cuda()
    d_packed = torch.from_numpy(packed.view(np.int32).copy()).cuda()
    d_out = torch.zeros(want.size + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(slots.size + 1, dtype=torch.int64, device="cuda")
    for shift in (0, 2, 6, 14):
        d_out.zero_()
        nf = C.c_uint64(0)
        rc = gpu.lib.tb200_gsmtap_pack(gpu.h, C.c_void_p(d_slots.data_ptr()), C.c_void_p(d_packed.data_ptr()), slots.size,
                                       C.c_void_p(d_out.data_ptr() + shift), want.size, C.c_void_p(d_off.data_ptr()), C.byref(nf), 1)
        assert rc == want.size, gpu.err()
        got = d_out.cpu().numpy()
        assert np.array_equal(got[shift:shift + want.size], want)
        assert not got[:shift].any() and not got[shift + want.size:].any()
        assert int(d_off[-1]) == want.size


@pytest.mark.gpu
def test_gsmtap_pack_gpu_million(gpu, orc):
    """10^6 device-resident slots of the bench workload: frame count and sizes from the slot flags, frames of
    sampled slots against the oracle's frames of the same records"""
    import torch
    n = 1_000_000
    cfg = T.GenCfg(seed=0x7E7A0002, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
    nbits = 510 * n
    d_bits = torch.empty(nbits + 64, dtype=torch.uint8, device="cuda")
    assert gpu.lib.tb200_gen_stream_dev(gpu.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 0) == 0
    ms = n + 16
    d_slots = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
    d_t1 = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
    d_pk = torch.zeros(ms * 9, dtype=torch.int32, device="cuda")
    gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_UNPACKED | T.OUT_PACKED)
    ns = gpu.lib.tb200_rx_stream_dev(gpu.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()),
                                     C.c_void_p(d_t1.data_ptr()), C.c_void_p(d_pk.data_ptr()), ms)
    assert ns == n - 1, gpu.err()
    nf = C.c_uint64(0)
    need = gpu.lib.tb200_gsmtap_pack(gpu.h, C.c_void_p(d_slots.data_ptr()), None, ns, None, 0, None, C.byref(nf), 1)
    slots = d_slots[:ns * 16].cpu().numpy().view(T.SLOT_DTYPE)
    kinds = slots["flags"] & 3
    a = (slots["flags"] & 4) != 0; b = (slots["flags"] & 8) != 0
    per_slot = np.where(kinds == 1, 18 + 24 * a + 32 * b, np.where(kinds == 2, 18 + 50 * a, np.where(kinds == 3, 18 + 32 * (a.astype(int) + b), 0)))
    assert need == int(per_slot.sum())
    d_out = torch.zeros(need, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(ns + 1, dtype=torch.int64, device="cuda")
    gpu.set_options(profile=1)
    try:
        rc = gpu.lib.tb200_gsmtap_pack(gpu.h, C.c_void_p(d_slots.data_ptr()), C.c_void_p(d_pk.data_ptr()), ns,
                                       C.c_void_p(d_out.data_ptr()), need, C.c_void_p(d_off.data_ptr()), C.byref(nf), 1)
        ms_leaf = gpu.timing().leaf_ms
    finally:
        gpu.set_options(profile=0)
    assert rc == need, gpu.err()
    off = d_off.cpu().numpy()
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(per_slot)]))
    out = d_out.cpu().numpy()
    rng = np.random.default_rng(9)
    for i in np.concatenate([[0, 1, 255, 256, ns - 1], rng.integers(0, ns, 200)]):
        i = int(i)
        t1 = d_t1[i * 288:(i + 1) * 288].cpu().numpy()
        rec = gpu.expand_records(slots[i:i + 1], t1[None, :])
        want, _ = orc.gsmtap_frames(rec)
        assert np.array_equal(out[off[i]:off[i + 1]], want), i
    algo = ns * (16 + 36) + need
    print(f"gsmtap_pack: {ns} slots, {need} bytes in {ms_leaf:.3f} ms = {algo / ms_leaf / 1e6:.0f} GB/s algorithmic")


def _random_slots(rng, n):
    """slot records of every kind / CRC / BNCH / time combination with random type-1 bits (no decode involved)"""
    slots = np.zeros(n, dtype=T.SLOT_DTYPE)
    kind = rng.integers(0, 4, n)
    slots["flags"] = kind | (rng.integers(0, 2, n) << 2) | (rng.integers(0, 2, n) << 3) | (rng.integers(0, 2, n) << 4)
    # NDB with SCH/F has no second CRC flag, and BNCH only exists on SYNC bursts
    slots["flags"] &= np.where(kind == 2, 0xff ^ 0x18, np.where(kind == 1, 0xff, 0xff ^ 0x10)).astype(np.uint8)
    slots["time"] = rng.integers(0, 5, n) | (rng.integers(0, 32, n) << 3) | (rng.integers(0, 64, n) << 8)   # tn = 0 included
    t1 = np.zeros((n, 288), dtype=np.uint8)
    nbits = np.array([0, 198, 282, 262])[kind]
    for i in range(n):
        t1[i, :nbits[i]] = rng.integers(0, 2, nbits[i])
    packed = np.packbits(t1, axis=1, bitorder="little").view("<u4").reshape(n, 9)
    return slots, t1, np.ascontiguousarray(packed)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 1000, 270_000])      # 270 000 slots: more tiles than scan threads
def test_gsmtap_pack_random_slots_emulated(emu, orc, ref, n):
    """every kind, flag and time value (tn = 0 gives timeslot 255 like the reference's uint8 tn - 1): kernels = oracle =
    the reference's tetra_gsmtap.c"""
    slots, t1, packed = _random_slots(np.random.default_rng(100 + n), n)
    rec = emu.expand_records(slots, t1)
    want, nf = orc.gsmtap_frames(rec)
    want_ref, nf_ref = ref.gsmtap_frames(rec)
    assert nf == nf_ref and np.array_equal(want, want_ref)
    got, off, ng = emu.gsmtap_pack(slots, packed)
    assert ng == nf and np.array_equal(got, want)
    assert off[-1] == want.size and np.all(np.diff(off.astype(np.int64)) <= 82)


@pytest.mark.gpu
def test_gsmtap_pack_random_slots_gpu(gpu, orc):
    slots, t1, packed = _random_slots(np.random.default_rng(7), 70_001)
    want, nf = orc.gsmtap_frames(gpu.expand_records(slots, t1))
    got, off, ng = gpu.gsmtap_pack(slots, packed)
    assert ng == nf and np.array_equal(got, want) and off[-1] == want.size
