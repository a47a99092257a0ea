"""GPU part of the input front ends: bit-packed and float32-symbol input through
the C ABI on a real GPU, against the oracle and against the byte-format path at full size."""
import ctypes as C

import numpy as np
import pytest

import tetra_testlib as T
from test_oracle import _stream

pytestmark = pytest.mark.gpu


def _want(orc, bits):
    orc.reset(); orc.feed(bits, 64)
    return orc.records(), orc.events()


def _dev(gpu, buf_t, nbits, ms):
    import torch
    ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
    dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
    ns = gpu.lib.tb200_rx_stream_dev(gpu.h, C.c_void_p(buf_t.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()),
                                     C.c_void_p(dt.data_ptr()), None, ms)
    assert ns >= 0, gpu.err()
    return ds[:ns * 16], dt[:ns * 288]


@pytest.mark.parametrize("lead_in", [333, 334, 0])
def test_packed_input(gpu, orc, lead_in):
    import torch
    bits, _ = _stream(orc, n=5000, random_cell=1, lead_in_bits=lead_in)
    want, ev = _want(orc, bits)
    pk = T.pack_bits(bits)
    gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_PACKED, output=T.OUT_UNPACKED | T.OUT_PACKED)
    try:
        for pieces in (0, 700):
            gpu.set_options(pipeline_slots=pieces)
            slots, t1, _ = gpu.rx_stream_host_raw(pk, bits.size)
            T.check_stream_against(want, ev, slots, gpu.expand_records(slots, t1))
        gpu.set_options(pipeline_slots=0)
        # device-resident, 16-byte aligned and only 4-byte aligned buffers
        for shift in (0, 4, 12):
            buf = torch.zeros(pk.size + 64, dtype=torch.uint8, device="cuda")
            view = buf[shift:shift + pk.size]
            view.copy_(torch.from_numpy(pk))
            ds, dt = _dev(gpu, view, bits.size, bits.size // 510 + 16)
            slots = ds.cpu().numpy().view(T.SLOT_DTYPE)
            T.check_stream_against(want, ev, slots, gpu.expand_records(slots, dt.cpu().numpy().reshape(-1, 288)))
    finally:
        gpu.set_options(input=T.IN_BYTES, pipeline_slots=0)


@pytest.mark.parametrize("lead_in", [333, 334])
def test_symbol_input(gpu, orc, lead_in):
    import torch
    bits, _ = _stream(orc, n=5000, random_cell=1, lead_in_bits=lead_in)
    bits = bits[:bits.size & ~1]
    rng = np.random.default_rng(11)
    sym = T.bits_to_symbols(bits, rng, edge_share=0.05)
    assert np.array_equal(orc.float_to_bits(sym), bits)
    want, ev = _want(orc, bits)
    gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_F32SYM, output=T.OUT_UNPACKED | T.OUT_PACKED)
    try:
        for pieces in (0, 900):
            gpu.set_options(pipeline_slots=pieces)
            slots, t1, _ = gpu.rx_stream_host_raw(sym, bits.size)
            T.check_stream_against(want, ev, slots, gpu.expand_records(slots, t1))
        gpu.set_options(pipeline_slots=0)
        for shift in (0, 1):                    # 16-byte aligned / only 4-byte aligned
            buf = torch.zeros(sym.size + 16, dtype=torch.float32, device="cuda")
            view = buf[shift:shift + sym.size]
            view.copy_(torch.from_numpy(sym))
            ds, dt = _dev(gpu, view, bits.size, bits.size // 510 + 16)
            slots = ds.cpu().numpy().view(T.SLOT_DTYPE)
            T.check_stream_against(want, ev, slots, gpu.expand_records(slots, dt.cpu().numpy().reshape(-1, 288)))
    finally:
        gpu.set_options(input=T.IN_BYTES, pipeline_slots=0)


def test_slicer_golden_on_device(gpu, orc):
    """the reference program's golden vector through the device slicer: embed the golden symbols as the
    lead-in of a stream, lock onto the bursts behind it and compare the search log (positions depend on
    every sliced lead-in bit count) plus a direct check that byte-format and symbol-format runs agree"""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slicer.npz"))
    gold_bits = np.unpackbits(z["bits"])[:2 * z["sym"].size]
    bits, _ = _stream(orc, n=300, random_cell=1, lead_in_bits=0)
    rng = np.random.default_rng(12)
    sym = np.concatenate([z["sym"], T.bits_to_symbols(bits, rng)])
    allbits = np.concatenate([gold_bits, bits])
    assert np.array_equal(orc.float_to_bits(sym), allbits)
    want, ev = _want(orc, allbits)
    gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_F32SYM)
    try:
        slots, t1, _ = gpu.rx_stream_host_raw(sym, allbits.size)
    finally:
        gpu.set_options(input=T.IN_BYTES)
    T.check_stream_against(want, ev, slots, gpu.expand_records(slots, t1))


def test_formats_agree_at_full_size(gpu, orc):
    """10^6 bursts: the packed and the symbol front end give exactly the byte-format result"""
    import torch
    n = 1_000_000
    cfg = T.GenCfg(seed=0x7E7A0006, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=334)
    nbits = 510 * n + 334
    d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
    assert gpu.lib.tb200_gen_stream_dev(gpu.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0, gpu.err()
    ms = n + 16
    gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, input=T.IN_BYTES, output=T.OUT_UNPACKED)
    s0, t0 = _dev(gpu, d, nbits, ms)
    assert s0.numel() // 16 == n - 1
    # pack on the device: stream bit i -> byte i>>3 bit i&7
    nb8 = (nbits + 7) // 8
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device="cuda")
    padded = torch.zeros(nb8 * 8, dtype=torch.uint8, device="cuda")
    padded[:nbits] = d[:nbits]
    pk = (padded.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)
    pk = torch.cat([pk, torch.zeros(64, dtype=torch.uint8, device="cuda")])
    # symbols on the device: 00 -> 1, 01 -> 3, 10 -> -1, 11 -> -3 (units of pi/4) plus noise inside the decision region
    pairs = padded[:nbits].view(-1, 2).to(torch.int64)
    code = pairs[:, 0] * 2 + pairs[:, 1]
    centre = torch.tensor([1.0, 3.0, -1.0, -3.0], device="cuda")[code]
    sym = (centre + (torch.rand(code.numel(), device="cuda") - 0.5) * 1.9).to(torch.float32).contiguous()
    del padded, pairs, code, centre
    try:
        gpu.set_options(input=T.IN_PACKED)
        s1, t1 = _dev(gpu, pk, nbits, ms)
        assert torch.equal(s0, s1) and torch.equal(t0, t1)
        gpu.set_options(input=T.IN_F32SYM)
        s2, t2 = _dev(gpu, sym, nbits, ms)
        assert torch.equal(s0, s2) and torch.equal(t0, t2)
    finally:
        gpu.set_options(input=T.IN_BYTES)
