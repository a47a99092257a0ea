"""BASELINE config 5 on real GPUs through the C driver (tb200_dist_*, NCCL inside libtetra_b200.so): ONE stream held by
rank 0, sharded over the ranks (grouped ncclSend / ncclRecv of bit-packed or one-bit-per-byte shards, or read in place
over NVLink), cell state carried by one all-gather, results rank-local.  Checked three ways: every run bit for bit
against the single-GPU decode of the whole stream, the per-rank digests against the single-GPU digest, and windows
that straddle the shard edges against the CPU oracle.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import tetra_testlib as T

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bursts, q, mode, wipe):
    try:
        _worker_body(rank, world, port, n_bursts, q, mode, wipe)
    except BaseException:
        import traceback
        q.put((rank, "error", traceback.format_exc()))
        raise


def _worker_body(rank, world, port, n_bursts, q, mode, wipe):
    import faulthandler
    import sys
    faulthandler.enable()
    faulthandler.dump_traceback_later(150, exit=True, file=sys.stderr)      # a rank stuck in a collective says where
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = T.B200(device=rank)
    cfg = T.GenCfg(seed=0x7E7A0005, sb_period=5, lead_sb=2, ndb2_per_256=64, ber_per_65536=655,
                   random_cell=1, lead_in_bits=333)
    nbits = 510 * n_bursts + 333
    buf = T.DevBuffer(g, nbits + 64)               # exportable memory (the peer mode maps it into the other ranks)
    full = buf.tensor(dev)
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n_bursts, C.c_void_p(full.data_ptr()), 1) == 0, g.err()
    for k in wipe:                                  # a wiped training sequence: lock is lost there (tetra_burst_sync.c:138-142)
        full[333 + 510 * k + 200:333 + 510 * k + 300] = 0
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_PACKED | T.OUT_UNPACKED)
    # reference result: the whole stream on this GPU alone
    ms = n_bursts + 16
    ds = torch.zeros(ms * 16, dtype=torch.uint8, device=dev)
    dt = torch.zeros(ms * 288, dtype=torch.uint8, device=dev)
    dp = torch.zeros(ms * 9, dtype=torch.int32, device=dev)
    ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(full.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()),
                                   C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), ms)
    assert ns > 0, (ns, g.err())
    losses = g.stats().lock_losses
    one_digest = g.slots_digest(ds.data_ptr(), dp.data_ptr(), 0, is_device=True, n=ns)
    # the sharded run through the C driver; the NCCL id goes out of band (here: a torch.distributed broadcast)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(T.Dist.get_id(g)), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    dd = T.Dist(g, rank, world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
    src = full
    if mode & 0xff == T.DIST_SCATTER and not (mode & T.DIST_PACK):
        pass
    mls = dd.max_local_slots(nbits)
    s_slots = torch.zeros(mls * 16, dtype=torch.uint8, device=dev)
    s_t1 = torch.zeros(mls * 288, dtype=torch.uint8, device=dev)
    s_pk = torch.zeros(mls * 9, dtype=torch.int32, device=dev)
    n, runs = dd.rx_stream(src.data_ptr() if rank == 0 else None, nbits, mode, s_slots.data_ptr(), s_t1.data_ptr(), s_pk.data_ptr(), mls)
    ok = True
    dig = 0
    for gs, l, c in runs:
        ok = ok and torch.equal(s_slots[l * 16:(l + c) * 16], ds[gs * 16:(gs + c) * 16]) \
            and torch.equal(s_t1[l * 288:(l + c) * 288], dt[gs * 288:(gs + c) * 288]) \
            and torch.equal(s_pk[l * 9:(l + c) * 9], dp[gs * 9:(gs + c) * 9])
        dig = (dig + g.slots_digest(s_slots.data_ptr() + 16 * l, s_pk.data_ptr() + 36 * l, gs, is_device=True, n=c)) & (2 ** 64 - 1)
    tot = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    codes = s_slots[:n * 16].view(torch.int32).view(-1, 4)[:, 1].unique().numel() if n else 0
    # oracle replay of a window that straddles this rank's first shard edge (config-4 style window: starts at an SB)
    oracle_ok = None
    if runs and not wipe:
        gs, l, c = runs[0]
        k_edge = gs + (1 if rank == 0 else 0)                     # slot index == burst index - 1
        k0 = max(5, (k_edge // 5) * 5 - 40) if rank else 5        # a burst that is a SYNC burst (sb_period 5)
        a = 333 + 510 * k0
        win = full[a:a + 510 * 120].cpu().numpy()
        orc = T.Oracle()
        orc.reset(); orc.feed(win, 64)
        want = orc.records()
        sel0, sel1 = k0 + 1 - 1, k0 + 118 - 1
        sl = ds[sel0 * 16:sel1 * 16].cpu().numpy().view(T.SLOT_DTYPE).copy()
        sl["slot_bit"] = (510 * (np.arange(sel0, sel1) + 1 - k0)).astype(np.uint32)
        got = g.expand_records(sl, dt[sel0 * 288:sel1 * 288].cpu().numpy().reshape(-1, 288))
        # the oracle starts cold at the window: it is in step with the stream's receiver from the first CRC-good SB1 on
        # (cell code and time, tetra_lower_mac.c:291-302); that must happen before the shard edge (burst 40 of the window)
        oracle_ok, msg = False, "never in step"
        for j in range(1, 40):
            w = want[(want["slot_bit"] >= 510 * j) & (want["slot_bit"] < 510 * 118)]
            h = got[(got["slot_bit"] >= 510 * j) & (got["slot_bit"] < 510 * 118)]
            same, msg = T.records_equal(w, h)
            if same:
                oracle_ok = True
                break
        # ... and the sharded result equals the single-GPU one there (checked above for the whole run)
    tm = dd.timing()
    q.put((rank, bool(ok), n, int(tot[0]), ns, codes, dig, one_digest, oracle_ok, int(losses), int(tm.segments)))
    dist.barrier()
    dd.close()
    dist.destroy_process_group()


CASES = [(2, T.DIST_SCATTER, False), (2, T.DIST_SCATTER | T.DIST_PACK, False), (2, T.DIST_PEER, False), (2, T.DIST_PEER | T.DIST_PACK, True),
         (2, T.DIST_SCATTER | T.DIST_PACK, True), (4, T.DIST_SCATTER | T.DIST_PACK, True), (8, T.DIST_SCATTER | T.DIST_PACK, False),
         (8, T.DIST_PEER | T.DIST_PACK, True)]


@pytest.mark.parametrize("world,mode,lose_lock", CASES)
def test_one_stream_sharded_over_gpus(world, mode, lose_lock):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n_bursts = 400_000
    # one wiped training sequence inside every shard: lock is lost `world` times
    wipe = [int((r + 0.37) * n_bursts / world) for r in range(world)] if lose_lock else []
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bursts, q, mode, wipe), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = []
        for _ in range(world):
            r = q.get(timeout=240)
            assert r[1] != "error", f"rank {r[0]}:\n{r[2]}"
            res.append(r)
        res.sort()
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:              # a rank that died leaves its peers inside a collective
            if p.is_alive():
                p.kill()
    digest_sum = 0
    for rank, ok, n, tot, ns, codes, dig, one_digest, oracle_ok, losses, segs in res:
        assert ok, f"rank {rank}: a run differs from the single-GPU decode"
        assert tot == ns                       # the runs together are every slot of the single-GPU run
        assert codes > min(1000, n // 8) or n == 0          # random cells: the carried cell state really mattered
        assert oracle_ok is None or oracle_ok, f"rank {rank}: window at the shard edge differs from the oracle"
        digest_sum = (digest_sum + dig) & (2 ** 64 - 1)
        if lose_lock:
            assert losses >= world and segs >= world + 1
    assert digest_sum == res[0][7]             # digests of the shards add up to the single-GPU digest
