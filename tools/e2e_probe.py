"""Host-buffer (e2e) timing probe: pinned host in/out through tb200_rx_stream_host, various piece sizes / outputs."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tetra_testlib as T

def main():
    n = 1_000_000
    g = T.B200()
    cfg = T.GenCfg(seed=0x7E7A0002, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
    nbits = 510 * n
    d_bits = torch.empty(nbits + 64, dtype=torch.uint8, device="cuda")
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 0) == 0
    ms = n + 16
    hb = g.lib.tb200_host_alloc(nbits); hs = g.lib.tb200_host_alloc(ms * 16); ht = g.lib.tb200_host_alloc(ms * 288); hp = g.lib.tb200_host_alloc(ms * 36)
    np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_uint8)), shape=(nbits,))[:] = d_bits[:nbits].cpu().numpy()
    # raw copy bandwidth for reference
    t = torch.empty(nbits, dtype=torch.uint8, device="cuda")
    src = torch.from_numpy(np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_uint8)), shape=(nbits,)))
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); t.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"plain pinned H2D of the stream: {dt*1e3:.2f} ms = {nbits/dt/1e9:.1f} GB/s")
    for out_mode, name in ((T.OUT_UNPACKED, "unpacked"), (T.OUT_PACKED, "packed"), (T.OUT_UNPACKED | T.OUT_PACKED, "both")):
        for P in (16384, 32768, 65536, 131072, 262144):
            g.set_options(output=out_mode, pipeline_slots=P, profile=0, viterbi=1)
            tp = ht if out_mode & T.OUT_UNPACKED else None
            pp = hp if out_mode & T.OUT_PACKED else None
            best = 1e9
            for it in range(5):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                ns = g.lib.tb200_rx_stream_host(g.h, hb, nbits, 3, hs, tp, pp, ms)
                dt = time.perf_counter() - t0
                assert ns == n - 1, g.err()
                if it: best = min(best, dt)
            print(f"out={name:8s} piece={P:7d}: {best*1e3:7.2f} ms -> {ns/best/1e6:7.1f} M bursts/s  (H2D {nbits/best/1e9:.1f} GB/s)")

def packed():
    """bit-packed input (64 B per burst): piece size sweep"""
    n = 1_000_000
    g = T.B200()
    cfg = T.GenCfg(seed=0x7E7A0002, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0)
    nbits = 510 * n
    d_bits = torch.empty(nbits + 64, dtype=torch.uint8, device="cuda")
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 0) == 0
    pk = np.packbits(d_bits[:nbits].cpu().numpy(), bitorder="little")
    ms = n + 16
    hb = g.lib.tb200_host_alloc(pk.size + 64); hs = g.lib.tb200_host_alloc(ms * 16); ht = g.lib.tb200_host_alloc(ms * 288); hp = g.lib.tb200_host_alloc(ms * 36)
    np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_uint8)), shape=(pk.size,))[:] = pk
    for out_mode, name in ((T.OUT_UNPACKED, "unpacked"), (T.OUT_PACKED, "packed")):
        for P in (0, 65536, 131072, 262144, 524288, 1048576):
            g.set_options(output=out_mode, pipeline_slots=P, profile=0, viterbi=1, input=T.IN_PACKED)
            tp = ht if out_mode & T.OUT_UNPACKED else None
            pp = hp if out_mode & T.OUT_PACKED else None
            best = 1e9
            for it in range(6):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                ns = g.lib.tb200_rx_stream_host(g.h, hb, nbits, 3, hs, tp, pp, ms)
                dt = time.perf_counter() - t0
                assert ns == n - 1, g.err()
                if it: best = min(best, dt)
            print(f"packed in, out={name:8s} piece={P:7d}: {best*1e3:7.2f} ms -> {ns/best/1e6:7.1f} M bursts/s")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "packed":
        packed()
        sys.exit(0)
    main()
