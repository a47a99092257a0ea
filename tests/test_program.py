"""The whole program: the reference's own `tetra-rx` (main(), real upper MAC, LLC, MLE, crypto; compiled in
place into oracle/_ref/tetra-rx with the libosmocore stand-ins) against the SAME program with its PHY +
lower MAC objects replaced by osmo-tetra_b200/host/tetra_shim.c + libtetra_b200 (src/Makefile:26 with the
change INTEGRATION.md shows).  Same input file, same command line: stdout must be byte-identical, upper-MAC
output and its interleaving with the PHY / lower-MAC lines included, and so must the traffic dump files.
CPU: the shim drives the SIMT-emulation build of the CUDA sources; marked gpu: the real library."""
import glob
import os
import subprocess
import tempfile

import numpy as np
import pytest

import tetra_testlib as T
from test_fuzz import make_case

REF_PROG = os.path.join(T.ROOT, "oracle", "_ref", "tetra-rx")
PROG_OBJ = os.path.join(T.ROOT, "oracle", "_ref", "prog")
REF_SRC = "/root/reference/src"
LOWER = ("phy/", "lower_mac/")

needs_ref = pytest.mark.skipif(not (os.path.exists(REF_PROG) and os.path.isdir(PROG_OBJ)),
                               reason="oracle/_ref/tetra-rx not built (no /root/reference here)")


def build_shim_program(lib_path, out):
    """tetra-rx.o + the reference's upper-layer objects + the shim instead of phy/*.o and lower_mac/*.o"""
    objs = [o for o in glob.glob(os.path.join(PROG_OBJ, "**", "*.o"), recursive=True)
            if not any(("/prog/" + d) in o for d in LOWER)]
    shim_o = out + "_shim.o"
    if os.path.isdir(REF_SRC):
        inc = ["-I" + REF_SRC, "-I" + os.path.join(T.ROOT, "oracle", "stubs"), "-I" + os.path.join(T.ROOT, "include")]
        subprocess.check_call(["gcc", "-O2", "-g", "-c", "-DTETRA_B200_SHIM_WRAP_READ"] + inc +
                              [os.path.join(T.ROOT, "osmo-tetra_b200", "host", "tetra_shim.c"), "-o", shim_o])
    else:
        # no reference headers here (GPU box): the object oracle/Makefile compiled in the build container
        shim_o = os.path.join(T.ROOT, "oracle", "_ref", "shim", "tetra_shim.o")
        if not os.path.exists(shim_o):
            pytest.skip("neither the reference headers nor a prebuilt shim object are present")
    subprocess.check_call(["gcc", "-o", out] + objs + [shim_o, lib_path, "-Wl,--wrap=read", "-Wl,-rpath," + os.path.dirname(lib_path)])
    return out


def _limits():
    # the reference's upper MAC can run away on damaged PDUs (it prints past the end of short messages):
    # bound what a child may write and how long it may run
    import resource
    resource.setrlimit(resource.RLIMIT_FSIZE, (64 << 20, 64 << 20))
    resource.setrlimit(resource.RLIMIT_CPU, (60, 60))


def run(prog, path, dumpdir, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = os.path.join(dumpdir, "..", os.path.basename(dumpdir) + ".stdout")
    with open(out, "wb") as fo:
        r = subprocess.run([prog, "-d", dumpdir, path], stdout=fo, stderr=subprocess.PIPE, env=e, preexec_fn=_limits, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    return open(out, "rb").read()


def compare(prog, bits, d, tag, dumps=True, env=None):
    path = os.path.join(d, tag + ".bits")
    np.ascontiguousarray(bits, dtype=np.uint8).tofile(path)
    ref_dir, our_dir = os.path.join(d, tag + "_ref"), os.path.join(d, tag + "_ours")
    os.makedirs(ref_dir); os.makedirs(our_dir)
    want = run(REF_PROG, path, ref_dir)
    got = run(prog, path, our_dir, env)
    assert want.count(b"TMV-UNITDATA.ind") > 20
    assert got == want, tag
    if dumps:
        assert sorted(os.listdir(ref_dir)) == sorted(os.listdir(our_dir))
        for f in os.listdir(ref_dir):
            assert open(os.path.join(ref_dir, f), "rb").read() == open(os.path.join(our_dir, f), "rb").read(), f
    return want


@needs_ref
def test_tetra_rx_with_shim_prints_what_tetra_rx_prints(orc):
    simt = T.build_simt()
    with tempfile.TemporaryDirectory() as d:
        prog = build_shim_program(simt, os.path.join(d, "tetra-rx-shim"))
        small = {"TETRA_B200_BATCH_BITS": "20000"}               # several GPU batches per file
        for seed in (1003, 1007, 1012):
            bits, _ = make_case(orc, seed, 80)                     # tetra-rx reads 64 bytes at a time
            compare(prog, bits, d, f"fuzz{seed}", dumps=False, env=small)
        # single-channel stream, BER 0: every block reaches the upper MAC; AACH values that mark traffic
        cfg = T.GenCfg(seed=0x7E7A0088, sb_period=7, lead_sb=2, ndb2_per_256=0, ber_per_65536=0, random_cell=0, lead_in_bits=200)
        bits = orc.gen_stream(cfg, 0, 150)
        out = compare(prog, bits, d, "clean", dumps=True, env=small)
        assert out.count(b"CRC COMP: 0x1d0f OK") > 150


@pytest.mark.gpu
@needs_ref
def test_tetra_rx_with_shim_on_gpu(gpu, orc):
    with tempfile.TemporaryDirectory() as d:
        prog = build_shim_program(T.PRODUCT_SO, os.path.join(d, "tetra-rx-shim"))
        small = {"TETRA_B200_BATCH_BITS": "20000"}
        # the same streams as on the CPU: the reference's upper MAC is only known to survive these random payloads
        for seed in (1003, 1007, 1012):
            bits, _ = make_case(orc, seed, 80)
            compare(prog, bits, d, f"fuzz{seed}", dumps=False, env=small)
        cfg = T.GenCfg(seed=0x7E7A0088, sb_period=7, lead_sb=2, ndb2_per_256=0, ber_per_65536=0, random_cell=0, lead_in_bits=200)
        compare(prog, orc.gen_stream(cfg, 0, 150), d, "clean", dumps=True)
