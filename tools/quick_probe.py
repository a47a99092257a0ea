"""Quick device-resident timing probe (not the bench): generate N bursts on the GPU, decode, wall-clock."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tetra_testlib as T

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    g = T.B200()
    cfg = T.GenCfg(seed=0x7E7A0002, sb_period=0, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
    nbits = 510 * n
    d_bits = torch.empty(nbits + 64, dtype=torch.uint8, device="cuda")
    t = time.time(); rc = g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 0); torch.cuda.synchronize()
    print("gen rc", rc, "%.3fs" % (time.time() - t))
    ms = n + 16
    d_slots = torch.empty(ms * 16, dtype=torch.uint8, device="cuda")
    d_t1 = torch.empty(ms * 288, dtype=torch.uint8, device="cuda")
    d_pk = torch.empty(ms * 9, dtype=torch.int32, device="cuda")
    for variant in (T.VITERBI_WARP, T.VITERBI_LANE):
        g.set_options(viterbi=variant, profile=1)
        for it in range(4):
            torch.cuda.synchronize(); t = time.time()
            ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()),
                                           C.c_void_p(d_t1.data_ptr()), C.c_void_p(d_pk.data_ptr()), ms)
            torch.cuda.synchronize(); dt = time.time() - t
            tm = g.timing()
            print(f"variant {variant} iter {it}: slots {ns} in {dt*1e3:.2f} ms -> {ns/dt/1e6:.2f} M bursts/s | dev total {tm.total_ms:.2f} classify {tm.classify_ms:.2f} scan {tm.scan_ms:.2f} decode {tm.decode_ms:.2f}", g.err() if ns < 0 else "")
        slots = np.frombuffer(d_slots.cpu().numpy().tobytes(), dtype=T.SLOT_DTYPE)[:ns]
        print(" kinds", np.unique(slots['flags'] & 3, return_counts=True), "crcA frac", ((slots['flags'] & 4) != 0).mean())

if __name__ == "__main__":
    main()
