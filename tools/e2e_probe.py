"""host-buffer path at one size with a given host_pack_threads (argv: bursts threads [steps]): ms per step"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import tetra_testlib as T
import bench
n = int(float(sys.argv[1])); threads = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = T.B200()
cfg = bench.gen_cfg(T, 0x7E7A0004, "config4")
lead = bench.SHAPES["config4"]["lead_in_bits"]
nbits = 510 * n + lead
d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0
ms = n + 16
hb = g.lib.tb200_host_alloc(nbits); hs = g.lib.tb200_host_alloc(ms * 16); ht = g.lib.tb200_host_alloc(ms * 288)
torch.from_numpy(np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_uint8)), shape=(nbits,))).copy_(d[:nbits])
del d
g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=0, host_pack_threads=threads)
def step():
    k = g.lib.tb200_rx_stream_host(g.h, hb, nbits, 3, hs, ht, None, ms)
    assert k > 0.99 * n, (k, g.err())
step()
t0 = time.perf_counter()
for _ in range(steps): step()
dt = (time.perf_counter() - t0) / steps
print("E2E bursts %d threads %d every %s: %.1f ms per step, %.4g bursts/s" % (n, threads, os.environ.get("TB200_HOST_PACK_EVERY", "1"), dt * 1e3, n / dt))
