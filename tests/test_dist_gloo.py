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
