"""Do the HBM-bound search kernel and the issue-bound decode kernel overlap when two receivers share one GPU?
Two contexts (own streams), each decoding its own 10^6-burst stream from its own host thread; aggregate
bursts/s against one context alone.  TB200_LANE_CTAS_PER_SM limits the decode CTAs per context."""
import ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tetra_testlib as T


def make(seed, n):
    g = T.B200()
    cfg = T.GenCfg(seed=seed, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
    nbits = 510 * n
    d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 0) == 0
    ms = n + 16
    ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
    dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=0)

    def step():
        ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()), C.c_void_p(dt.data_ptr()), None, ms)
        assert ns == n - 1, (ns, g.err())
    return step


def run(steps_fns, k):
    def loop(fn):
        for _ in range(k):
            fn()
    for fn in steps_fns:
        fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=loop, args=(fn,)) for fn in steps_fns]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


n, k = 1_000_000, 40
nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fns = [make(0x7E7A0002 + i, n) for i in range(nctx)]
dt = run(fns, k)
print(f"contexts {nctx} ctas/sm {os.environ.get('TB200_LANE_CTAS_PER_SM', 'max')}: {nctx * (n - 1) * k / dt:.4g} bursts/s aggregate, {dt / k * 1e3:.3f} ms per round of {nctx} steps")
