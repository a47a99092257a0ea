import ctypes as C, os, sys
sys.path.insert(0, "tests")
import torch, tetra_testlib as T
g = T.B200(); n = 1_000_000
cfg = T.GenCfg(seed=0x7E7A0002, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
nbits = 510 * n
d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 0)
ms = n + 16
ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda"); dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
g.set_options(chunk_bits=64, viterbi=1, output=1, pipeline_slots=0, profile=0)
for i in range(6):
    if i == 5: os.environ["TB200_TRACE"] = "1"
    g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()), C.c_void_p(dt.data_ptr()), None, ms)
