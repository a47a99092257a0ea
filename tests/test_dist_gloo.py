"""The N > 1 path of bench.py on CPU: two ranks over gloo, each decoding its own self-contained
stream (SIMT-emulation build of the library), timings reduced with MAX, counts summed."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tetra_testlib as T

sys.path.insert(0, T.ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    g = T.B200(emulate=True)
    orc = T.Oracle()
    cfg = bench.gen_cfg(T, bench.rank_seed(rank))
    bits = orc.gen_stream(cfg, 0, 70)
    slots, t1, _ = g.rx_stream_host(bits)
    got = g.expand_records(slots, t1)
    orc.reset(); orc.feed(bits, 64)
    ok, msg = T.records_equal(orc.records(), got)
    # every rank reports a fake device time; the slowest must win on all ranks
    my_ms = 10.0 + rank
    red = bench.reduce_max(dist, [my_ms, float(slots.size)], "cpu")
    total = torch.tensor([slots.size], dtype=torch.int64)
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, ok, msg, red, int(total[0]), int(slots.size), int(bits[:2000].sum())))
    dist.destroy_process_group()


def test_two_ranks_independent_shards():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    for rank, ok, msg, red, total, mine, chk in res:
        assert ok, (rank, msg)
        assert red[0] == 11.0                      # max over ranks, seen by every rank
        assert total == sum(r[5] for r in res)     # whole-job count = sum of the shards
    assert res[0][6] != res[1][6]                  # the ranks really decoded different streams


def shard_plan(n_slots_total, world):
    """contiguous slot ranges per rank (the last ranks get the remainder)"""
    per = (n_slots_total + world - 1) // world
    return [(min(r * per, n_slots_total), min((r + 1) * per, n_slots_total)) for r in range(world)]


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = T.B200(emulate=True)
    orc = T.Oracle()
    cfg = T.GenCfg(seed=77, sb_period=9, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)
    nb = 150
    bits = np.ascontiguousarray(orc.gen_stream(cfg, 0, nb))          # every rank can regenerate the stream
    n_end = bits.size
    # rank 0 acquires lock on the head of the stream and tells the others where slot 0 is
    meta = torch.zeros(3, dtype=torch.int64)
    if rank == 0:
        ok, a0, cmin = g.find_lock(bits.ctypes.data, n_end)
        meta[:] = torch.tensor([int(ok), a0, cmin])
    dist.broadcast(meta, 0)
    ok, a0, cmin = int(meta[0]), int(meta[1]), int(meta[2])
    assert ok
    n_total = (n_end - a0) // 510                # with 64-bit reads every full slot of this stream gets processed
    k0, k1 = shard_plan(n_total, world)[rank]
    n = k1 - k0
    # the rank only looks at its own slice of the stream (+ look-ahead halo)
    lo = a0 + 510 * k0
    hi = min(n_end, a0 + 510 * k1 + 4096 + 64)
    shard = np.ascontiguousarray(bits[lo:hi])
    s = g.shard_pass1(shard.ctypes.data, lo, shard.size, lo, cmin + k0, n_end, n)
    # the one exchange step of the path: all-gather of the 32-byte summaries
    mine = torch.frombuffer(bytearray(bytes(s)), dtype=torch.uint8)
    gathered = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, mine)
    summaries = [T.ShardSummary.from_buffer_copy(bytes(t.numpy().tobytes())) for t in gathered]
    assert all(x.first_unlock == 0xffffffff for x in summaries)
    carry = g.shard_carry_in(summaries, rank)
    slots = np.zeros(max(n, 1), dtype=T.SLOT_DTYPE)
    t1 = np.zeros((max(n, 1), 288), dtype=np.uint8)
    got_n = g.shard_pass2(carry, slots.ctypes.data, t1.ctypes.data)
    assert got_n == n
    rec = g.expand_records(slots[:n], t1[:n])
    orc.reset(); orc.feed(bits, 64)
    want = orc.records()
    mine_want = want[(want["slot_bit"] >= lo) & (want["slot_bit"] < a0 + 510 * k1)]
    okr, msg = T.records_equal(mine_want, rec)
    q.put((rank, okr, msg, n, int(want.size), int(rec.size)))
    dist.barrier()
    dist.destroy_process_group()


def test_one_stream_sharded_over_two_ranks():
    """config 5 shape on CPU: one stream, contiguous slot shards, cell state carried by an all-gather"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, msg, n, want_total, got in res:
        assert ok, (rank, msg)
    assert sum(r[5] for r in res) == res[0][4]       # the shards together give every record of the stream


# ---------------------------------------------------------------- the C driver (tb200_dist_*) over gloo

def _dist_streams(orc):
    """(name, bits, mode) cases for the C driver: plain, packed on rank 0, and a training sequence wiped in every shard"""
    cfg = T.GenCfg(seed=78, sb_period=7, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)
    nb = 170
    bits = np.ascontiguousarray(orc.gen_stream(cfg, 0, nb))
    wiped = bits.copy()
    for k in (40, 120):                      # normal bursts in the first and in the second rank's shard: lock is lost twice
        while orc.gen_kind(cfg, k) == 1:
            k += 1
        wiped[333 + 510 * k + 244:333 + 510 * k + 266] = 0
    sparse = T.GenCfg(seed=79, sb_period=120, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=17)
    far = np.ascontiguousarray(orc.gen_stream(sparse, 0, nb))     # the second shard holds no SYNC burst at all
    return [("plain", bits, T.DIST_SCATTER), ("packed", bits, T.DIST_SCATTER | T.DIST_PACK),
            ("lock lost in both shards", wiped, T.DIST_SCATTER), ("lock lost, packed", wiped, T.DIST_SCATTER | T.DIST_PACK),
            ("no SB in the second shard", far, T.DIST_SCATTER)]


def _dist_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = T.B200(emulate=True)
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED | T.OUT_PACKED, pipeline_slots=0)
    ops = T.gloo_dist_ops(dist)
    dd = T.Dist(g, rank, world, ops=ops)
    orc = T.Oracle()
    out = []
    for name, bits, mode in _dist_streams(orc):
        ms = bits.size // 510 + 16
        slots = np.zeros(ms, dtype=T.SLOT_DTYPE)
        t1 = np.zeros((ms, 288), dtype=np.uint8)
        pk = np.zeros((ms, 9), dtype=np.uint32)
        n, runs = dd.rx_stream(bits.ctypes.data if rank == 0 else None, bits.size, mode, slots.ctypes.data, t1.ctypes.data,
                               pk.ctypes.data, ms)
        dig = sum(g.slots_digest(slots[l:l + c], pk[l:l + c], k_base=gs) for gs, l, c in runs) & (2 ** 64 - 1)
        cy = g.carry()
        out.append((name, n, runs, slots[:n].copy(), t1[:n].copy(), pk[:n].copy(), dig, (cy.state, cy.scramb_init, cy.tn, cy.fn, cy.mn),
                    dd.timing().segments))
    q.put((rank, out))
    dist.barrier()
    dd.close()
    dist.destroy_process_group()


def test_c_driver_two_ranks(orc):
    """tb200_dist_rx_stream (the C sharding driver) with gloo plumbing on two CPU ranks: the runs of both ranks, put
    in global slot order, are exactly what ONE receiver delivers - records and search log against the oracle -
    also when lock is lost inside a shard, and the per-rank digests add up to the digest of the single run"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    emu = T.B200(emulate=True)
    emu.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED | T.OUT_PACKED, pipeline_slots=0)
    for ci, (name, bits, mode) in enumerate(_dist_streams(orc)):
        orc.reset(); orc.feed(bits, 64)
        want, ev = orc.records(), orc.events()
        pieces = []
        for rank in range(world):
            _, n, runs, slots, t1, pk, dig, cy, segs = res[rank][ci]
            assert sum(c for _, _, c in runs) == n, name
            for gs, l, c in runs:
                pieces.append((gs, slots[l:l + c], t1[l:l + c], pk[l:l + c]))
        pieces.sort(key=lambda x: x[0])
        pos = 0
        for gs, s, _, _ in pieces:               # the runs tile the global slot order without gaps
            assert gs == pos, (name, gs, pos)
            pos += s.size
        slots = np.concatenate([p[1] for p in pieces]); t1 = np.concatenate([p[2] for p in pieces]); pk = np.concatenate([p[3] for p in pieces])
        T.check_stream_against(want, ev, slots, emu.expand_records(slots, t1))
        one_s, one_t1, one_pk = emu.rx_stream_host(bits)
        assert np.array_equal(one_s, slots) and np.array_equal(one_pk, pk), name
        total = sum(res[r][ci][6] for r in range(world)) & (2 ** 64 - 1)
        assert total == emu.slots_digest(one_s, one_pk) == T.slots_digest_host(one_s, one_pk), name
        # rank 0 holds the receiver state a single receiver ends with
        c = emu.carry()
        assert res[0][ci][7] == (c.state, c.scramb_init, c.tn, c.fn, c.mn), name
        if "lock lost" in name:
            assert res[0][ci][8] >= 3, name          # 1 + two losses
