"""options.host_pack_threads on the GPU (an opt-in host-side feature, off by default): in a file of its own that sorts last."""
import numpy as np
import pytest

import tetra_testlib as T
from test_gpu import _check, _gen_on_gpu, _stream

pytestmark = pytest.mark.gpu


def test_host_pack_threads_gpu(gpu, orc):
    """options.host_pack_threads: host threads pack the caller's bytes before the copy - same output as the plain path over many
    pipeline pieces (ramp included), and the oracle's records on a stream with a lead-in"""
    bits, _ = _stream(orc, n=9000, random_cell=1, ber_per_65536=1200)
    try:
        _check(gpu, orc, bits, viterbi=T.VITERBI_LANE, pipeline_slots=1024, host_pack_threads=4)
        cfg = T.GenCfg(seed=0x7E7A0055, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)
        d, nbits = _gen_on_gpu(gpu, cfg, 600_000)
        big = d[:nbits].cpu().numpy()
        gpu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, host_pack_threads=0)
        s0, t0, p0 = gpu.rx_stream_host(big)
        gpu.set_options(host_pack_threads=6)
        s1, t1, p1 = gpu.rx_stream_host(big)
        assert s0.size > 599_000 and np.array_equal(s0, s1) and np.array_equal(t0, t1) and np.array_equal(p0, p1)
    finally:
        gpu.set_options(host_pack_threads=0, pipeline_slots=0)
