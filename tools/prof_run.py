"""One workload shape, a few device-resident steps with serial passes: the command ncu wraps for the per-kernel captures.
usage: python tools/prof_run.py <config2|config3|config4> <bursts> [steps]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
import tetra_testlib as T
import bench

shape, n = sys.argv[1], int(float(sys.argv[2]))
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = T.B200()
cfg = bench.gen_cfg(T, 0x7E7A0004, shape)
lead = bench.SHAPES[shape]["lead_in_bits"]
nbits = 510 * n + lead
d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0, g.err()
ms = n + 16
ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=1, serial_passes=1)
for _ in range(steps):
    ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()), C.c_void_p(dt.data_ptr()), None, ms)
    assert ns > 0.99 * n, (ns, g.err())
t = g.timing()
kinds = bench.kinds_of(torch, ds, ns)
print("PROF %s slots %d kinds %s decode_ms %.4f prepare_ms %.4f trellis_ms %.4f search_ms %.4f classify_ms %.4f scan_ms %.4f total_ms %.4f" %
      (shape, ns, kinds, t.decode_ms, t.prepare_ms, t.trellis_ms, t.search_ms, t.classify_ms, t.scan_ms, t.total_ms))
