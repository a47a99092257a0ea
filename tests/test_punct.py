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
