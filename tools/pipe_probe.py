"""Device-resident throughput at a size with many pieces: config-2 and config-4 shapes (A/B of the piece pipeline)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tetra_testlib as T

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_000_000
    g = T.B200()
    res = []
    for name, kw in (("cfg2", dict(sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)),
                     ("cfg4", dict(sb_period=2, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333))):
        cfg = T.GenCfg(seed=0x7E7A0004, **kw)
        nbits = 510 * n + kw["lead_in_bits"]
        d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
        assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0, g.err()
        ms = n + 16
        ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
        dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
        g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=0)
        def step():
            ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()), C.c_void_p(dt.data_ptr()), None, ms)
            assert ns > 0.99 * n, (ns, g.err())
            return ns
        for _ in range(3):
            step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        k = 10
        for _ in range(k):
            ns = step()
        torch.cuda.synchronize(); dt_s = (time.perf_counter() - t0) / k
        res.append("%s %.4g/s (%.3f ms)" % (name, ns / dt_s, dt_s * 1e3))
        if os.environ.get("PROBE_PROF"):
            g.set_options(profile=1, serial_passes=1)
            step(); step()
            t = g.timing(); per = 1e6 / ns
            res.append("[per 1e6: search %.4f sb1 %.4f scan %.4f decode %.4f = prep %.4f + trellis %.4f + fin %.4f]" % (
                t.search_ms * per, (t.classify_ms - t.search_ms) * per, t.scan_ms * per, t.decode_ms * per,
                t.prepare_ms * per, t.trellis_ms * per, (t.decode_ms - t.prepare_ms - t.trellis_ms) * per))
            g.set_options(profile=0, serial_passes=0)
        del d, ds, dt
    print("  ".join(res))

main()
