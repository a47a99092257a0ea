"""Committed golden vectors (made from the reference's own code by tests/golden/make_golden.py)
against the oracle and against the CUDA sources run under the CPU SIMT emulator."""
import numpy as np
import pytest

import tetra_testlib as T


@pytest.mark.parametrize("name", ["config1_sb.npz", "mixed_noisy.npz"])
def test_oracle_matches_golden_stream(orc, name):
    bits, rec, ev = T.load_golden_stream(name)
    orc.reset(); orc.feed(bits, 64)
    ok, msg = T.records_equal(rec, orc.records())
    assert ok, msg
    assert np.array_equal(ev, orc.events())


@pytest.mark.parametrize("name", ["sb1", "ndb", "schf"])
def test_oracle_matches_golden_blocks(orc, name):
    bt, t5, codes, t1, ok = T.load_golden_blocks(name)
    for i in range(t5.shape[0]):
        orc.reset(); orc.set_cell(int(codes[i])); orc.tp_sap(bt, 1, t5[i])
        r = orc.records()[0]
        assert np.array_equal(r["type1"][:t1.shape[1]], t1[i]) and r["crc_ok"] == ok[i], i


def test_config1_shape():
    """config 1 facts the survey predicts: 999 decoded bursts, SB1 good, SB2 CRC wrong (generator quirk)"""
    bits, rec, ev = T.load_golden_stream("config1_sb.npz")
    assert bits.size == 510000 and rec.size == 2997
    assert rec["crc_ok"][rec["lchan"] == T.LC_BSCH].all()
    assert not rec["crc_ok"][(rec["type1_len"] == 124)].any()


def test_emulated_kernels_match_golden_mixed(emu):
    bits, rec, ev = T.load_golden_stream("mixed_noisy.npz")
    n = 333 + 510 * 130            # covers the wiped training sequence at burst 100 and the re-lock
    import tetra_testlib as TT
    ref_rec = rec[rec["slot_bit"] + 510 <= n]
    emu.set_options(viterbi=T.VITERBI_WARP, chunk_bits=64, pipeline_slots=0)
    slots, t1, _ = emu.rx_stream_host(bits[:n])
    got = emu.expand_records(slots, t1)
    # cutting the stream short changes nothing before the cut except the very last window
    keep = got["slot_bit"] + 510 + 64 <= n
    ok, msg = T.records_equal(ref_rec[:keep.sum()], got[keep])
    assert ok, msg
    assert emu.stats().lock_losses == 1


@pytest.mark.parametrize("name", ["sb1", "ndb", "schf"])
@pytest.mark.parametrize("variant", [T.VITERBI_WARP, T.VITERBI_LANE])
def test_emulated_kernels_match_golden_blocks(emu, name, variant):
    bt, t5, codes, t1, ok = T.load_golden_blocks(name)
    emu.set_options(viterbi=variant)
    out, crc = emu.decode_blocks(bt, t5[:24], codes[:24])
    assert np.array_equal(out, t1[:24])
    assert np.array_equal(crc, ok[:24])
