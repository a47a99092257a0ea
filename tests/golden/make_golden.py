"""Generates the committed golden fixtures from the REFERENCE's own code (oracle/_ref, compiled in
place from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py

  config1_sb.npz    BASELINE config 1: 1 000 SYNC bursts built by the reference's conv_enc_test
                    generator (build_sb, conv_enc_test.c:198-305, rand() replaced by a fixed-seed
                    PRNG), fed to the reference receiver in 64-byte reads; its records.
                    Known generator quirks are frozen in: SB2 encodes sb_type2 with an over-read
                    (:273) and SB2/AACH are not scrambled (:284,:296), see SURVEY.md 8d.
  mixed_noisy.npz   600 mixed SB / NDB bursts at BER 1e-2 with one wiped training sequence (lock
                    loss + re-acquisition), one early false NORM hit and a SYNC hit at the wrong
                    offset; the reference receiver's records and search log.
  blocks_noisy.npz  64 noisy coded blocks per block type with the reference lower MAC's output.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import tetra_testlib as T  # noqa: E402


def pack_records(rec):
    return dict(slot_bit=rec["slot_bit"], lchan=rec["lchan"], crc_ok=rec["crc_ok"], blk_num=rec["blk_num"],
                tn=rec["tn"], fn=rec["fn"], mn=rec["mn"], type1_len=rec["type1_len"],
                scrambling_code=rec["scrambling_code"], type1=np.packbits(rec["type1"], axis=1))


def main():
    ref, orc = T.Ref(), T.Oracle()

    # ---- config 1
    state = 0x7E7A0001
    bursts = []
    out = np.zeros(510, np.uint8)
    for _ in range(1000):
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        r = (state >> 33) & 0x7fffffff                         # rand()-like 31-bit value
        assert ref.lib.ref_conv_enc_test_sb(C.c_uint32(r), out.ctypes.data_as(C.c_void_p)) == 0
        bursts.append(out.copy())
    bits = np.concatenate(bursts)
    ref.reset(); ref.feed(bits, 64)
    rec, ev = ref.records(), ref.events()
    np.savez_compressed(os.path.join(HERE, "config1_sb.npz"), bits=np.packbits(bits), n_bits=bits.size,
                        events=ev, **pack_records(rec))
    print("config1: bursts 1000 records", rec.size, "crc ok", int(rec["crc_ok"].sum()))

    # ---- mixed noisy
    cfg = T.GenCfg(seed=0x7E7A0003, sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655,
                   random_cell=1, lead_in_bits=333)
    bits = orc.gen_stream(cfg, 0, 600)
    seq = {"y": "11000001100111001110100111000001100111", "p": "0111101001000011011110"}
    o = lambda k: 333 + 510 * k
    bits[o(100) + 244:o(100) + 266] = 0                                   # training sequence wiped
    bits[o(300) + 30:o(300) + 52] = T.bits_from_str(seq["p"])             # early false NORM_2 hit
    bits[o(450) + 100:o(450) + 138] = T.bits_from_str(seq["y"])           # SYNC at the wrong offset
    ref.reset(); ref.feed(bits, 64)
    rec, ev = ref.records(), ref.events()
    np.savez_compressed(os.path.join(HERE, "mixed_noisy.npz"), bits=np.packbits(bits), n_bits=bits.size,
                        events=ev, **pack_records(rec))
    print("mixed: records", rec.size, "events", ev.size, "crc ok", int(rec["crc_ok"].sum()))

    # ---- noisy blocks through the reference lower MAC (tp_sap_udata_ind)
    rng = np.random.default_rng(0x7E7A)
    blocks = {}
    for bt, name in ((T.T_SB1, "sb1"), (T.T_NDB, "ndb"), (T.T_SCH_F, "schf")):
        K, N, T1, a = T.BLK[bt]
        t5s, codes, t1s, oks = [], [], [], []
        for i in range(64):
            t1 = rng.integers(0, 2, T1).astype(np.uint8)
            t2 = np.zeros(N, np.uint8); t2[:T1] = t1
            crc = (~ref.crc16(t2[:T1])) & 0xffff
            t2[T1:T1 + 16] = [(crc >> (15 - b)) & 1 for b in range(16)]
            # the reference keeps its cell code in a static; it can only be set through a SYNC PDU,
            # so non-SB1 blocks use the code of a cell announced by a synthetic SB1 first
            mcc, mnc, cc = (int(rng.integers(0, 1 << 10)), int(rng.integers(0, 1 << 14)), int(rng.integers(0, 64)))
            code = 3 if bt == T.T_SB1 else ref.scramb_get_init(mcc, mnc, cc)
            t5 = ref.scramb_bits(code, ref.interleave(K, a, ref.punct_2_3(ref.conv_encode(t2), K)))
            ber = [0.0, 0.01, 0.03, 0.06][i % 4]
            t5 ^= (rng.random(K) < ber).astype(np.uint8)
            ref.reset()
            if bt != T.T_SB1:
                pdu = np.zeros(80, np.uint8)
                f = lambda v, n: [(v >> (n - 1 - j)) & 1 for j in range(n)]
                pdu[:60] = f(0, 4) + f(cc, 6) + f(0, 2) + f(1, 5) + f(1, 6) + f(0, 8) + f(mcc, 10) + f(mnc, 14) + f(0, 5)
                c2 = (~ref.crc16(pdu[:60])) & 0xffff
                pdu[60:76] = f(c2, 16)
                sb1 = ref.scramb_bits(3, ref.interleave(120, 11, ref.punct_2_3(ref.conv_encode(pdu), 120)))
                ref.tp_sap(T.T_SB1, 1, sb1)
                assert ref.scramb_init() == code
            ref.tp_sap(bt, 1, t5)
            r = ref.records()[-1]
            t5s.append(t5); codes.append(code); t1s.append(r["type1"][:T1].copy()); oks.append(r["crc_ok"])
        blocks[name + "_type5"] = np.packbits(np.array(t5s), axis=1)
        blocks[name + "_code"] = np.array(codes, np.uint32)
        blocks[name + "_type1"] = np.packbits(np.array(t1s), axis=1)
        blocks[name + "_crc_ok"] = np.array(oks, np.uint8)
        print(name, "crc ok", int(np.sum(oks)), "of 64")
    np.savez_compressed(os.path.join(HERE, "blocks_noisy.npz"), **blocks)


if __name__ == "__main__":
    main()
