"""Input front ends in front of the hot path: the symbol slicer of
float_to_bits.c and bit-packed input.  CPU part: the oracle restatement against the reference's own
float_to_bits program and its golden vector; the CUDA kernels under the SIMT emulator against the
oracle.  The GPU part of the same checks lives in test_gpu_frontend.py."""
import os
import tempfile

import numpy as np
import pytest

import tetra_testlib as T
from test_oracle import _stream

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slicer.npz")


def test_slicer_oracle_matches_golden(orc):
    z = np.load(GOLD)
    want = np.unpackbits(z["bits"])[:2 * z["sym"].size]
    assert np.array_equal(orc.float_to_bits(z["sym"]), want)


def test_slicer_oracle_matches_reference_program(orc):
    if not os.path.exists(T.REF_FLOAT_TO_BITS):
        pytest.skip("oracle/_ref/float_to_bits not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    sym = (rng.standard_normal(100_000) * 3).astype(np.float32)
    sym[::97] = 2.0; sym[1::97] = 0.0; sym[2::97] = -2.0; sym[3::97] = np.nan
    with tempfile.TemporaryDirectory() as d:
        want = T.ref_float_to_bits(sym, d)
    assert np.array_equal(orc.float_to_bits(sym), want)


def test_symbols_round_trip(orc):
    rng = np.random.default_rng(6)
    bits = rng.integers(0, 2, 20_000, dtype=np.uint8)
    assert np.array_equal(orc.float_to_bits(T.bits_to_symbols(bits, rng, edge_share=0.2)), bits)


def _ref_records(orc, bits):
    orc.reset(); orc.feed(bits, 64)
    return orc.records(), orc.events()


@pytest.mark.parametrize("lead_in", [333, 334, 0, 77])
def test_packed_input_emulated(emu, orc, lead_in):
    """bit-packed stream in, same records as the reference chain on the unpacked stream (lock loss included)"""
    bits, _ = _stream(orc, n=260, random_cell=1, lead_in_bits=lead_in)
    want, ev = _ref_records(orc, bits)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_PACKED)
    try:
        slots, t1, _ = emu.rx_stream_host_raw(T.pack_bits(bits), bits.size)
        T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
        emu.set_options(pipeline_slots=100)                    # several pieces: unaligned piece starts
        slots, t1, _ = emu.rx_stream_host_raw(T.pack_bits(bits), bits.size)
        T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
    finally:
        emu.set_options(input=T.IN_BYTES, pipeline_slots=0)


@pytest.mark.parametrize("lead_in", [333, 334])
def test_symbol_input_emulated(emu, orc, lead_in):
    """float32 symbols in: sliced on the 'device' like float_to_bits, then the usual chain"""
    bits, _ = _stream(orc, n=200, random_cell=1, lead_in_bits=lead_in)
    bits = bits[:bits.size & ~1]
    rng = np.random.default_rng(7)
    sym = T.bits_to_symbols(bits, rng)
    assert np.array_equal(orc.float_to_bits(sym), bits)
    want, ev = _ref_records(orc, bits)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_F32SYM)
    try:
        slots, t1, _ = emu.rx_stream_host_raw(sym, bits.size)
        T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
        emu.set_options(pipeline_slots=70)
        slots, t1, _ = emu.rx_stream_host_raw(sym, bits.size)
        T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
    finally:
        emu.set_options(input=T.IN_BYTES, pipeline_slots=0)


def _dev_call(lib, part, n_bits, flags):
    """tb200_rx_stream_dev on a numpy buffer (emulation build)"""
    ms = n_bits // 510 + 16
    slots = np.zeros(ms, dtype=T.SLOT_DTYPE); t1 = np.zeros((ms, 288), dtype=np.uint8); pk = np.zeros((ms, 9), dtype=np.uint32)
    k = lib.rx_stream_dev_raw(part.ctypes.data, n_bits, flags, slots.ctypes.data, t1.ctypes.data, pk.ctypes.data, ms)
    return slots[:k], t1[:k], pk[:k]


def _continued(lib, orc, fmt, n_bursts, seed, dev=False):
    """a packed / symbol stream handed over in several calls (128-bit boundaries) equals the reference chain on the whole stream"""
    rng = np.random.default_rng(seed)
    bits, cfg = _stream(orc, n=n_bursts, random_cell=1, sb_period=8, lead_in_bits=int(rng.integers(0, 300)))
    bits = bits[:bits.size & ~1].copy()
    k = n_bursts // 2
    while orc.gen_kind(cfg, k) == 1:
        k += 1
    bits[cfg.lead_in_bits + 510 * k + 244:cfg.lead_in_bits + 510 * k + 266] = 0        # and lock is lost on the way
    want, ev = _ref_records(orc, bits)
    buf = T.pack_bits(bits) if fmt == T.IN_PACKED else T.bits_to_symbols(bits, rng)
    per_bit = 1 / 8 if fmt == T.IN_PACKED else 2                              # buffer elements (bytes / floats) per 1 stream bit... floats: 1 per 2 bits
    lib.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=fmt)
    try:
        pos, outs = 0, []
        while pos < bits.size:
            n = min(bits.size - pos, 128 * int(rng.integers(1, 60)))
            last = pos + n >= bits.size
            part = buf[pos // 8:(pos + n + 7) // 8] if fmt == T.IN_PACKED else buf[pos // 2:(pos + n + 1) // 2]
            flags = (T.TB200_FRESH if pos == 0 else 0) | (T.TB200_FINAL if last else 0)
            if dev and len(outs) % 3 != 2:            # (emulation: host addresses are the "device" pointers; every third call the host way)
                outs.append(_dev_call(lib, np.ascontiguousarray(part), n, flags))
            else:
                outs.append(lib.rx_stream_host_raw(np.ascontiguousarray(part), n, flags=flags))
            pos += n
        slots = np.concatenate([o[0] for o in outs]); t1 = np.concatenate([o[1] for o in outs])
        T.check_stream_against(want, ev, slots, lib.expand_records(slots, t1))
        assert len(outs) > 5
    finally:
        lib.set_options(input=T.IN_BYTES, pipeline_slots=0)


@pytest.mark.parametrize("fmt", [T.IN_PACKED, T.IN_F32SYM])
def test_packed_and_symbol_streams_continue(emu, orc, fmt):
    _continued(emu, orc, fmt, 60, 41 + fmt)


@pytest.mark.parametrize("fmt", [T.IN_PACKED, T.IN_F32SYM])
def test_packed_and_symbol_streams_continue_dev(emu, orc, fmt):
    """the same through tb200_rx_stream_dev (kept bits on the device, staged head), host calls mixed in"""
    _continued(emu, orc, fmt, 60, 43 + fmt, dev=True)
    _continued(emu, orc, fmt, 150, 47 + fmt, dev=True)


def test_format_rules(emu):
    emu.set_options(input=T.IN_PACKED, viterbi=T.VITERBI_LANE)
    try:
        with pytest.raises(RuntimeError):          # a packed stream continues on 128-bit boundaries only
            emu.rx_stream_host_raw(np.zeros(64, np.uint8), 500, flags=T.TB200_FRESH)
        with pytest.raises(ValueError):
            emu.set_options(viterbi=T.VITERBI_WARP)
    finally:
        emu.set_options(input=T.IN_BYTES, viterbi=T.VITERBI_LANE)


def _drifting_symbols(rng, n):
    """demodulator output with a frequency offset: the symbols ride on a slowly moving mean, plus the odd outlier and NaN"""
    bits = rng.integers(0, 2, 2 * n).astype(np.uint8)
    sym = T.bits_to_symbols(bits, rng, edge_share=0.0)
    drift = (0.9 * np.sin(np.arange(n) / 9000.0) + 0.3).astype(np.float32)
    sym = (sym + drift).astype(np.float32)
    sym[rng.integers(0, n, n // 500)] = np.float32(7.5)          # beyond +-5: the tracker skips them (float_to_bits.c:141)
    sym[rng.integers(0, n, 5)] = np.float32(np.nan)
    return sym


def test_afc_oracle_matches_reference_program(orc):
    """float_to_bits -a: the oracle's tracker, rounding step by step like the C program, against the program itself"""
    if not os.path.exists(T.REF_FLOAT_TO_BITS):
        pytest.skip("oracle/_ref/float_to_bits not built (no /root/reference here)")
    rng = np.random.default_rng(17)
    sym = _drifting_symbols(rng, 300_000)
    with tempfile.TemporaryDirectory() as d:
        for args, kw in ((("-a",), {}), (("-a", "-f", "0.003"), dict(filter_val=0.003)), (("-a", "-f", "0.0005", "-F", "0.25"), dict(filter_val=0.0005, filter_goal=0.25))):
            want = T.ref_float_to_bits(sym, d, args)
            got, _ = orc.float_to_bits_afc(sym, **kw)
            assert np.array_equal(got, want), args
            assert not np.array_equal(got, orc.float_to_bits(sym))          # and it matters on this input
    # the tracker's state carries over: two halves = the whole
    a, st = orc.float_to_bits_afc(sym[:123_457])
    b, _ = orc.float_to_bits_afc(sym[123_457:], state=st)
    assert np.array_equal(np.concatenate([a, b]), orc.float_to_bits_afc(sym)[0])


def _afc_leaf(lib, orc, n, seeds):
    for seed, kw in seeds:
        rng = np.random.default_rng(seed)
        sym = _drifting_symbols(rng, n)
        want, wst = orc.float_to_bits_afc(sym, **kw)
        got, gst, redone = lib.float_to_bits(sym, afc=True, **kw)
        assert np.array_equal(got, want), (seed, kw)
        assert np.float32(gst).tobytes() == np.float32(wst).tobytes()
        # piece by piece with the carried state
        cut = int(rng.integers(1, n - 1)) & ~15
        a, st, _ = lib.float_to_bits(sym[:cut], afc=True, **kw)
        b, st2, _ = lib.float_to_bits(sym[cut:], afc=True, state=st, **kw)
        assert np.array_equal(np.concatenate([a, b]), want) and np.float32(st2).tobytes() == np.float32(wst).tobytes()
    plain, _, _ = lib.float_to_bits(sym, afc=False)
    assert np.array_equal(plain, orc.float_to_bits(sym))


def test_afc_leaf_emulated(emu, orc):
    """tb200_float_to_bits with the pseudo-AFC: chunks run speculatively from a warm-up and are verified against their
    predecessors, so the result is the serial program's bit for bit - also with a filter constant so small that the warm-up
    cannot converge and chunks have to be redone"""
    _afc_leaf(emu, orc, 60_000, [(1, {}), (2, dict(filter_val=0.01)), (3, dict(filter_val=0.002, filter_goal=0.3)), (4, dict(filter_val=0.5))])


def test_symbol_stream_with_afc_emulated(emu, orc):
    """TB200_IN_F32SYM with options.afc = 1 == float_to_bits -a | tetra-rx: the whole chain on a drifting symbol stream"""
    rng = np.random.default_rng(21)
    bits, cfg = _stream(orc, n=120, random_cell=1, lead_in_bits=334)
    bits = bits[:bits.size & ~127]
    sym = T.bits_to_symbols(bits, rng, edge_share=0.0)
    sym = (sym + (0.8 * np.sin(np.arange(sym.size) / 3000.0) + 0.4).astype(np.float32)).astype(np.float32)      # a wandering carrier offset
    kw = dict(filter_val=0.004)
    sliced, _ = orc.float_to_bits_afc(sym, **kw)
    assert not np.array_equal(orc.float_to_bits(sym), sliced)
    want, ev = _ref_records(orc, sliced)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_F32SYM, afc=1, afc_filter_val=0.004, afc_filter_goal=0.0)
    try:
        slots, t1, _ = emu.rx_stream_host_raw(sym, bits.size)
        T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
        # the stream in three calls: the tracker's state and the receiver's carry over
        outs, pos = [], 0
        for n in (128 * 40, 128 * 111, bits.size - 128 * 151):
            flags = (T.TB200_FRESH if pos == 0 else 0) | (T.TB200_FINAL if pos + n == bits.size else 0)
            outs.append(emu.rx_stream_host_raw(np.ascontiguousarray(sym[pos // 2:(pos + n) // 2]), n, flags=flags))
            pos += n
        s3 = np.concatenate([o[0] for o in outs]); t3 = np.concatenate([o[1] for o in outs])
        T.check_stream_against(want, ev, s3, emu.expand_records(s3, t3))
        # and through the device-resident call
        outs, pos = [], 0
        for n in (128 * 40, 128 * 111, bits.size - 128 * 151):
            flags = (T.TB200_FRESH if pos == 0 else 0) | (T.TB200_FINAL if pos + n == bits.size else 0)
            outs.append(_dev_call(emu, np.ascontiguousarray(sym[pos // 2:(pos + n) // 2]), n, flags))
            pos += n
        s4 = np.concatenate([o[0] for o in outs]); t4 = np.concatenate([o[1] for o in outs])
        T.check_stream_against(want, ev, s4, emu.expand_records(s4, t4))
    finally:
        emu.set_options(input=T.IN_BYTES, afc=0)


@pytest.mark.gpu
def test_afc_gpu(gpu, orc):
    _afc_leaf(gpu, orc, 3_000_000, [(5, {}), (6, dict(filter_val=0.003, filter_goal=-0.2))])
    # a filter constant far too small for the warm-up to settle inside the stream: chunks are redone, the bits still exact
    rng = np.random.default_rng(8)
    sym = _drifting_symbols(rng, 400_000)
    want, _ = orc.float_to_bits_afc(sym, filter_val=2e-6)
    got, _, redone = gpu.float_to_bits(sym, afc=True, filter_val=2e-6)
    assert np.array_equal(got, want)
