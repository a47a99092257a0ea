"""tetra-rx-b200 (osmo-tetra_b200/host/tetra_rx_b200.c), the file-in / text-out receiver: its stdout must be
byte-identical to what the reference's PHY + lower MAC print for the same file (found SYNC ..., BURST,
BNCH FOLLOWS, CRC COMP: 0x.... OK/WRONG, <blk> <time> type1: ..., TMB-SAP SYNC ...), including the CRC
register values of damaged blocks.  The reference text comes from oracle/_ref (the reference's own code
compiled in place) run in a child process with its stdout captured."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import tetra_testlib as T
from test_fuzz import make_case

HOST = os.path.join(T.ROOT, "osmo-tetra_b200", "host")

REF_CHILD = r'''
import sys, ctypes as C, numpy as np
sys.path.insert(0, sys.argv[1])
import tetra_testlib as T
ref = T.Ref()
bits = np.fromfile(sys.argv[2], dtype=np.uint8)
ref.reset()
ref.lib.ref_set_recording(0)
ref.lib.ref_feed(bits.ctypes.data_as(C.c_void_p), bits.size, int(sys.argv[3]), 0)
'''


def reference_stdout(path, chunk):
    return subprocess.run([sys.executable, "-c", REF_CHILD, os.path.join(T.ROOT, "tests"), path, str(chunk)],
                          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout


def build_cli(lib_path, out):
    subprocess.check_call(["gcc", "-O1", "-g", "-I" + os.path.join(T.ROOT, "include"), os.path.join(HOST, "tetra_rx_b200.c"),
                           lib_path, "-Wl,-rpath," + os.path.dirname(lib_path), "-o", out])
    return out


def _cases(orc, seeds, n_bursts):
    for seed in seeds:
        bits, chunk = make_case(orc, seed, n_bursts)
        yield seed, bits, chunk


@pytest.mark.skipif(not T.have_ref(), reason="reference build (oracle/_ref) not present")
def test_cli_text_matches_reference_emulated(orc):
    simt = T.build_simt()
    with tempfile.TemporaryDirectory() as d:
        cli = build_cli(simt, os.path.join(d, "tetra-rx-b200"))
        for seed, bits, chunk in _cases(orc, (1003, 1007, 1011, 1016), 70):
            path = os.path.join(d, f"s{seed}.bits")
            bits.tofile(path)
            want = reference_stdout(path, chunk)
            got = subprocess.run([cli, "-c", str(chunk), path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            assert want.count(b"CRC COMP") > 50
            assert got == want, (seed, chunk)
        # the other encodings of the same stream give the same text
        seed, bits, chunk = next(_cases(orc, (1003,), 70))
        bits = bits[:bits.size & ~1]
        path = os.path.join(d, "even.bits")
        bits.tofile(path)
        want = reference_stdout(path, chunk)
        T.pack_bits(bits).tofile(os.path.join(d, "p.bin"))
        got = subprocess.run([cli, "-f", "packed", "-c", str(chunk), os.path.join(d, "p.bin")], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, check=True).stdout
        # the packed file is padded to 16 bytes: the padding bits are zeros behind the stream and may add BURST-less tail only
        assert got.startswith(want[:len(want) - 200]) and got.count(b"CRC COMP") == want.count(b"CRC COMP")
        T.bits_to_symbols(bits, np.random.default_rng(3)).tofile(os.path.join(d, "s.f32"))
        got = subprocess.run([cli, "-f", "f32", "-c", str(chunk), os.path.join(d, "s.f32")], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, check=True).stdout
        assert got == want


@pytest.mark.gpu
@pytest.mark.skipif(not T.have_ref(), reason="reference build (oracle/_ref) not present")
def test_cli_text_matches_reference_gpu(gpu, orc):
    with tempfile.TemporaryDirectory() as d:
        cli = build_cli(T.PRODUCT_SO, os.path.join(d, "tetra-rx-b200"))
        for seed, bits, chunk in _cases(orc, (5001, 5002, 5005, 5009), 1200):
            path = os.path.join(d, f"s{seed}.bits")
            bits.tofile(path)
            want = reference_stdout(path, chunk)
            got = subprocess.run([cli, "-c", str(chunk), path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            assert want.count(b"CRC COMP") > 1000
            assert got == want, (seed, chunk)


@pytest.mark.gpu
@pytest.mark.skipif(not T.have_ref(), reason="reference build (oracle/_ref) not present")
def test_cli_multi_gpu_text_matches_reference(gpu, orc):
    """tetra-rx-b200 -G N: the C program decodes ONE stream on N GPUs (a host thread per GPU around tb200_dist_rx_stream,
    NCCL inside the library) and prints the reference's text, lock losses inside shards included.  N = 1 on a one-GPU box
    (a communicator of one rank: the same driver), N = all GPUs of a bigger box as well."""
    import torch
    worlds = sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()})
    with tempfile.TemporaryDirectory() as d:
        cli = build_cli(T.PRODUCT_SO, os.path.join(d, "tetra-rx-b200"))
        for seed, bits, chunk in _cases(orc, (5001, 5005), 1600):
            bits = bits.copy()
            for k in (400, 1100):                     # wiped training sequences: lock is lost twice
                bits[510 * k + 150:510 * k + 660] = 0
            path = os.path.join(d, f"m{seed}.bits")
            bits.tofile(path)
            want = reference_stdout(path, chunk)
            assert want.count(b"found SYNC") >= 3 and want.count(b"CRC COMP") > 1000
            for w in worlds:
                got = subprocess.run([cli, "-c", str(chunk), "-G", str(w), path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                     check=True, timeout=120).stdout
                assert got == want, (seed, chunk, w)
        packed = os.path.join(d, "p.bin")
        bits = bits[:bits.size & ~127]
        bits.tofile(path)
        T.pack_bits(bits).tofile(packed)
        want = reference_stdout(path, chunk)
        got = subprocess.run([cli, "-f", "packed", "-c", str(chunk), "-G", str(worlds[-1]), packed], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, check=True, timeout=120).stdout
        assert got == want


def _pcap_payloads(path):
    """(timestamps in us, UDP payloads) of a LINKTYPE_RAW pcap written by tetra-rx-b200 -p; checks the headers"""
    raw = open(path, "rb").read()
    magic, ver, _, _, snap, link = np.frombuffer(raw[:24], dtype="<u4")
    assert magic == 0xa1b2c3d4 and ver == (2 | 4 << 16) and link == 101
    p, ts, out = 24, [], []
    while p < len(raw):
        sec, usec, incl, orig = np.frombuffer(raw[p:p + 16], dtype="<u4")
        pkt = raw[p + 16:p + 16 + incl]
        assert incl == orig and pkt[0] == 0x45 and pkt[9] == 17 and int.from_bytes(pkt[2:4], "big") == incl
        words = np.frombuffer(pkt[:20], dtype=">u2").astype(np.uint32).sum()
        while words >> 16:
            words = (words & 0xffff) + (words >> 16)
        assert words == 0xffff                                        # IPv4 header checksum
        assert int.from_bytes(pkt[22:24], "big") == 4729 and int.from_bytes(pkt[24:26], "big") == incl - 20
        ts.append(int(sec) * 1_000_000 + int(usec)); out.append(pkt[28:])
        p += 16 + incl
    return ts, out


def _check_pcap(cli, orc, d, seed, n_bursts):
    bits, chunk = make_case(orc, seed, n_bursts)
    path = os.path.join(d, f"g{seed}.bits")
    bits.tofile(path)
    pc = os.path.join(d, f"g{seed}.pcap")
    subprocess.run([cli, "-c", str(chunk), "-p", pc, path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    orc.reset(); orc.feed(bits, chunk)
    rec = orc.records()
    good = rec[rec["crc_ok"] != 0]
    ts, frames = _pcap_payloads(pc)
    assert len(frames) == good.size > 20
    want, _ = orc.gsmtap_frames(rec)
    assert b"".join(frames) == want.tobytes()
    assert ts == sorted(ts) and ts[-1] > 0


def test_cli_pcap_emulated(orc):
    """-p: one IPv4/UDP packet to port 4729 per CRC-good block, payload = the reference's GSMTAP frame"""
    simt = T.build_simt()
    with tempfile.TemporaryDirectory() as d:
        cli = build_cli(simt, os.path.join(d, "tetra-rx-b200"))
        for seed in (2001, 2004):
            _check_pcap(cli, orc, d, seed, 90)


@pytest.mark.gpu
def test_cli_pcap_gpu(gpu, orc):
    with tempfile.TemporaryDirectory() as d:
        cli = build_cli(T.PRODUCT_SO, os.path.join(d, "tetra-rx-b200"))
        _check_pcap(cli, orc, d, 2101, 1500)
