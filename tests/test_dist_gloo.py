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
