import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tetra_testlib as T
g = T.B200(); orc = T.Oracle()
n = 10_000_000
cfg = T.GenCfg(seed=0x7E7A0003, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)
nbits = 510 * n + 333
d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0
# check generator at a few far positions against the CPU twin
for k in (0, 5_000_000, 7_499_998, 7_500_001, 9_999_999):
    got = d[333 + 510 * k: 333 + 510 * (k + 1)].cpu().numpy()
    want = orc.gen_stream(cfg, k, 1, lead_in=False)
    print("burst", k, "gen equal", np.array_equal(got, want), got[:12], want[:12])
ms = n + 16
ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda"); dp = torch.zeros(ms * 9, dtype=torch.int32, device="cuda")
g.set_options(viterbi=1, output=T.OUT_PACKED)
ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()), None, C.c_void_p(dp.data_ptr()), ms)
print("ns", ns, g.err())
st = g.stats(); c = g.carry()
print("stats slots", st.slots, "lock losses", st.lock_losses, "acq", st.lock_acquisitions, "launches", st.kernel_launches)
print("carry state", c.state, "calls", c.calls, "buf_start", c.buf_start_bit, "bits_in_buf", c.bits_in_buf, "stream_bits", c.stream_bits)
slots = ds[:ns * 16].cpu().numpy().view(T.SLOT_DTYPE)
print(slots[-3:])
