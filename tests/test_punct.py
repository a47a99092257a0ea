"""tetra_rcpc_depunct as a leaf operator (tb200_rcpc_depunct), all seven puncturers of the reference
(tetra_conv_enc.c:128-198), on the nine (type-2 length, type-3 length, mother rate, puncturer) tuples the
reference's own tetra_punct_test() walks (tetra_conv_enc.c:257-267).  Checked against the reference function
compiled in place (oracle/_ref) and through the reference's own loop-back property: puncturing a mother
buffer and de-puncturing the result gives back the mother buffer wherever something was sent."""
import ctypes as C

import numpy as np
import pytest

import tetra_testlib as T

# type2_len, type3_len, mother rate, enum tetra_rcpc_puncturer           (tetra_conv_enc.c:257-267)
TUPLES = [(80, 120, 4, 0), (292, 432, 4, 2), (148, 432, 4, 3), (144, 216, 4, 0), (112, 168, 4, 0), (288, 432, 4, 0),
          (112, 168, 3, 4), (72, 162, 3, 5), (38, 80, 3, 6)]


def _ref_depunct(ref, pu, type3, mother_len):
    out = np.full((type3.shape[0], mother_len), 0xff, dtype=np.uint8)
    for b in range(type3.shape[0]):
        row = np.ascontiguousarray(type3[b])
        o = np.full(mother_len + 64, 0xff, dtype=np.uint8)          # slack: the reference does not bound k
        assert ref.lib.tetra_rcpc_depunct(pu, T._ptr(row), row.size, T._ptr(o)) == 0
        assert (o[mother_len:] == 0xff).all()
        out[b] = o[:mother_len]
    return out


def _ref_punct(ref, pu, mother, type3_len):
    out = np.zeros(type3_len, dtype=np.uint8)
    m = np.ascontiguousarray(mother)
    assert ref.lib.get_punctured_rate(pu, T._ptr(m), type3_len, T._ptr(out)) == 0
    return out


def _run(g, ref, t2, t3, rate, pu, n):
    rng = np.random.default_rng(1000 * pu + t3)
    type3 = rng.integers(0, 2, (n, t3), dtype=np.uint8)
    want = _ref_depunct(ref, pu, type3, t2 * rate)
    got = g.rcpc_depunct(pu, type3, t2 * rate)
    assert np.array_equal(got, want)
    assert int((got != 0xff).sum()) == n * t3                     # every type-3 bit lands on its own mother position
    # the reference's loop-back (test_one_punct, tetra_conv_enc.c:283-320): bytes 0, 1, 2, ... as the mother buffer
    mother = (np.arange(t2 * rate) & 0x7f).astype(np.uint8)
    p3 = _ref_punct(ref, pu, mother, t3)
    back = g.rcpc_depunct(pu, p3[None, :], t2 * rate)[0]
    sent = back != 0xff
    assert sent.sum() == t3 and np.array_equal(back[sent], mother[sent])


@pytest.mark.parametrize("t2,t3,rate,pu", TUPLES)
def test_depunct_emulated(emu, ref, t2, t3, rate, pu):
    _run(emu, ref, t2, t3, rate, pu, 5)


def test_depunct_errors(emu):
    t3 = np.zeros((1, 120), dtype=np.uint8)
    with pytest.raises(RuntimeError):
        emu.rcpc_depunct(7, t3, 320)                               # ARRAY_SIZE(tetra_puncts) = 7: -EINVAL in the reference
    out = emu.rcpc_depunct(0, np.zeros((0, 120), dtype=np.uint8), 320)
    assert out.shape == (0, 320)


@pytest.mark.gpu
@pytest.mark.parametrize("t2,t3,rate,pu", TUPLES)
def test_depunct_gpu(gpu, ref, t2, t3, rate, pu):
    _run(gpu, ref, t2, t3, rate, pu, 300)


# ---- every rate-1/4 RCPC rate end to end: de-puncture + the Viterbi wrapper as leaves

RATE4 = [t for t in TUPLES if t[2] == 4]


def _roundtrip(g, ref, orc, t2, t3, pu, n, ber):
    """type-2 bits -> reference encoder + puncturer -> noise -> tb200_rcpc_depunct -> tb200_viterbi_decode, against the
    reference's own tetra_rcpc_depunct + viterbi_dec_sb1_wrapper (on the osmo_conv_decode stand-in) block by block"""
    rng = np.random.default_rng(77 * pu + t3)
    type2 = np.zeros((n, t2), dtype=np.uint8)
    type2[:, :t2 - 4] = rng.integers(0, 2, (n, t2 - 4))            # four tail bits bring the encoder back to state 0
    type3 = np.zeros((n, t3), dtype=np.uint8)
    for b in range(n):
        type3[b] = _ref_punct(ref, pu, ref.conv_encode(type2[b]), t3)
    type3 ^= (rng.random((n, t3)) < ber).astype(np.uint8)
    mother = g.rcpc_depunct(pu, type3, 4 * t2)
    got = g.viterbi_decode(mother, t2)
    want_m = _ref_depunct(ref, pu, type3, 4 * t2)
    assert np.array_equal(mother, want_m)
    for b in range(n):
        assert np.array_equal(got[b], ref.viterbi(want_m[b], t2)), (pu, b)
        assert np.array_equal(got[b], orc.viterbi(want_m[b], t2)), (pu, b)
    if ber == 0:
        assert np.array_equal(got, type2)


@pytest.mark.parametrize("t2,t3,rate,pu", RATE4)
def test_all_rates_decode_end_to_end_emulated(emu, ref, orc, t2, t3, rate, pu):
    _roundtrip(emu, ref, orc, t2, t3, pu, 6, 0.0)
    _roundtrip(emu, ref, orc, t2, t3, pu, 12, 0.04)


@pytest.mark.gpu
@pytest.mark.parametrize("t2,t3,rate,pu", RATE4)
def test_all_rates_decode_end_to_end_gpu(gpu, ref, orc, t2, t3, rate, pu):
    _roundtrip(gpu, ref, orc, t2, t3, pu, 40, 0.0)
    _roundtrip(gpu, ref, orc, t2, t3, pu, 200, 0.04)
    # the other tie rule moves oracle, stand-in and kernel together here too
    try:
        orc.set_tie(T.TIE_HIGH_PRED); ref.set_tie(T.TIE_HIGH_PRED); gpu.set_options(viterbi_tie=T.TIE_HIGH_PRED)
        _roundtrip(gpu, ref, orc, t2, t3, pu, 60, 0.04)
    finally:
        orc.set_tie(T.TIE_LOW_PRED); ref.set_tie(T.TIE_LOW_PRED); gpu.set_options(viterbi_tie=T.TIE_LOW_PRED)
