"""Differential fuzzing of the whole receive path: random downlink streams with random damage
(wiped training sequences -> lock loss and re-acquisition, training sequences planted at random
offsets -> early false hits, the pre-filter blind spot at offsets 0..20, misplaced SYNC sequences),
random read sizes and lead-ins.  The CUDA sources under the SIMT emulator (CPU) and, marked gpu, the
real kernels must reproduce the reference chain's records AND its search log bit for bit.  The oracle
restatement is the comparand here; it is itself pinned on the reference's compiled code by
test_oracle.py (the same mutations are part of test_fsm_* there)."""
import numpy as np
import pytest

import tetra_testlib as T
from tetra_testlib import bits_from_str as B
from test_oracle import SEQS


def make_case(orc, seed, n_bursts):
    rng = np.random.default_rng(seed)
    cfg = T.GenCfg(seed=int(rng.integers(1, 2**31)), sb_period=int(rng.integers(3, 20)), lead_sb=2,
                   ndb2_per_256=int(rng.choice([0, 64, 128])), ber_per_65536=int(rng.choice([0, 655, 2000])),
                   random_cell=int(rng.integers(0, 2)), lead_in_bits=int(rng.integers(0, 700)))
    bits = orc.gen_stream(cfg, 0, n_bursts).copy()
    lead = cfg.lead_in_bits
    n_mut = int(rng.integers(0, 6))
    for _ in range(n_mut):
        k = int(rng.integers(2, n_bursts - 1))
        a = lead + 510 * k
        kind = int(rng.integers(0, 5))
        if kind == 0:                                   # wipe the training sequence area: lock loss
            bits[a + 200:a + 290] = rng.integers(0, 2, 90)
            bits[a + 214:a + 252] = 0
        elif kind == 1:                                 # early false hit somewhere in the payload
            seq = B(SEQS[int(rng.choice([T.TS_NORM_1, T.TS_NORM_2]))])
            o = int(rng.integers(21, 200))
            bits[a + o:a + o + seq.size] = seq
        elif kind == 2:                                 # blind spot: a sequence at offsets 0..24 of the slot
            seq = B(SEQS[int(rng.choice([T.TS_NORM_1, T.TS_NORM_2, T.TS_SYNC]))])
            o = int(rng.integers(0, 25))
            bits[a + o:a + o + seq.size] = seq
            if o:
                bits[a + o - 1] = int(rng.integers(0, 2))
        elif kind == 3:                                 # SYNC sequence at a wrong offset: lock loss
            seq = B(SEQS[T.TS_SYNC])
            o = int(rng.integers(0, 460))
            bits[a + o:a + o + seq.size] = seq
        else:                                           # sequence only in the look-ahead of the previous slot
            bits[a + 244:a + 266] = rng.integers(0, 2, 22)
            seq = B(SEQS[T.TS_NORM_1])
            o = int(rng.integers(0, 60))
            bits[a + 510 + o:a + 510 + o + seq.size] = seq
    chunk = int(rng.choice([64, 64, 64, 1, 7, 100, 296, int(rng.integers(1, 297))]))
    cut = int(rng.integers(0, 200))
    if cut:
        bits = bits[:bits.size - cut]                   # ragged end
    return bits, chunk


def run_case(dev, orc, bits, chunk, **opts):
    orc.reset(); orc.feed(bits, chunk)
    want, ev = orc.records(), orc.events()
    dev.set_options(chunk_bits=chunk, **opts)
    slots, t1, _ = dev.rx_stream_host(bits)
    T.check_stream_against(want, ev, slots, dev.expand_records(slots, t1))
    c = dev.carry()
    assert c.state == orc.rx_state() and c.scramb_init == orc.scramb_init()


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_emulated(emu, orc, seed):
    bits, chunk = make_case(orc, 1000 + seed, 90)
    run_case(emu, orc, bits, chunk, viterbi=T.VITERBI_LANE, pipeline_slots=int(np.random.default_rng(seed).choice([0, 33, 64])))
    emu.set_options(chunk_bits=64, pipeline_slots=0)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_emulated_warp_form(emu, orc, seed):
    bits, chunk = make_case(orc, 2000 + seed, 50)
    run_case(emu, orc, bits, chunk, viterbi=T.VITERBI_WARP, pipeline_slots=0)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(40))
def test_fuzz_gpu(gpu, orc, seed):
    bits, chunk = make_case(orc, 5000 + seed, 1500)
    run_case(gpu, orc, bits, chunk, viterbi=T.VITERBI_LANE, pipeline_slots=int(np.random.default_rng(seed).choice([0, 200, 777])))
    gpu.set_options(chunk_bits=64, pipeline_slots=0)
