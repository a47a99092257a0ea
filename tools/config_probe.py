"""Device-resident timing of the other BASELINE configs' shapes (mixed kinds, per-burst random cells)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tetra_testlib as T

def run(name, n, **kw):
    g = T.B200()
    cfg = T.GenCfg(seed=0x7E7A0003, **kw)
    nbits = 510 * n + cfg.lead_in_bits
    d_bits = torch.empty(nbits + 64, dtype=torch.uint8, device="cuda")
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 1) == 0
    ms = n + 16
    d_slots = torch.empty(ms * 16, dtype=torch.uint8, device="cuda"); d_t1 = torch.empty(ms * 288, dtype=torch.uint8, device="cuda")
    g.set_options(viterbi=1, profile=1, output=T.OUT_UNPACKED)
    best = None
    for it in range(4):
        ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()), C.c_void_p(d_t1.data_ptr()), None, ms)
        assert ns > 0, g.err()
        tm = g.timing()
        if best is None or tm.total_ms < best[0]: best = (tm.total_ms, tm.classify_ms, tm.scan_ms, tm.decode_ms)
    slots = np.frombuffer(d_slots.cpu().numpy().tobytes(), dtype=T.SLOT_DTYPE)[:ns]
    kinds = np.bincount(slots['flags'] & 3, minlength=4)
    print(f"{name}: {ns} slots, dev {best[0]:.2f} ms (classify {best[1]:.2f} scan {best[2]:.2f} decode {best[3]:.2f}) -> {ns/best[0]/1e3:.0f} M bursts/s; kinds none/SB/F/2 = {kinds.tolist()}, lock losses {g.stats().lock_losses}, crcA {((slots['flags']&4)!=0).mean():.3f}")

run("config2 SCH/F           ", 1_000_000, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
run("config3 mixed, 1 cell   ", 1_000_000, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=0, lead_in_bits=333)
run("config3 mixed, rnd cells", 1_000_000, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)
run("config4 SB every other  ", 1_000_000, sb_period=2, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=1, lead_in_bits=0)
run("config4 all SB          ", 1_000_000, sb_period=1, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=1, lead_in_bits=0)
run("all two-block NDB       ", 1_000_000, sb_period=64, lead_sb=2, ndb2_per_256=256, ber_per_65536=655, random_cell=0, lead_in_bits=0)
