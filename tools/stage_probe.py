"""stand-alone descramble+deinterleave stage: correctness vs oracle on a sample + timing"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tetra_testlib as T
g = T.B200(); orc = T.Oracle()
g.set_options(profile=1)
for nblk in (1000, 33, 4_000_000):
    d5 = torch.randint(0, 2, (nblk * 432,), dtype=torch.uint8, device="cuda")
    d3 = torch.zeros_like(d5)
    dc = torch.randint(0, 2 ** 31 - 1, (nblk,), dtype=torch.int32, device="cuda")
    ms = []
    for _ in range(5):
        rc = g.lib.tb200_descramble_deinterleave(g.h, C.c_void_p(d5.data_ptr()), C.c_void_p(d3.data_ptr()), C.c_void_p(dc.data_ptr()), nblk, 432, 103, 1)
        assert rc == 0, g.err()
        ms.append(g.timing().leaf_ms)
    h5 = d5.cpu().numpy().reshape(nblk, 432); h3 = d3.cpu().numpy().reshape(nblk, 432); hc = dc.cpu().numpy().view(np.uint32)
    idx = list(range(min(nblk, 40))) + [nblk - 1, nblk // 2]
    bad = sum(not np.array_equal(h3[i], orc.deinterleave(432, 103, orc.scramb_bits(int(hc[i]), h5[i]))) for i in idx)
    t = min(ms[1:])
    print(f"nblk {nblk}: mismatches {bad}/{len(idx)}, {t:.3f} ms, {2*432*nblk/t/1e6:.0f} GB/s")
