"""BASELINE config 5 on real GPUs: ONE stream held by rank 0, scattered over the ranks with NCCL
send/recv, cell state carried by an all-gather of the 32-byte shard summaries, results rank-local.
Every rank then checks its shard bit for bit against a single-GPU decode of the whole stream
(the device generator is deterministic, so each rank can rebuild the stream for the check).
Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
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


def _worker(rank, world, port, n_bursts, q, peer=False):
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
    full = torch.zeros(nbits + 64, dtype=torch.uint8, device=dev)
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n_bursts, C.c_void_p(full.data_ptr()), 1) == 0, g.err()
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_PACKED | T.OUT_UNPACKED)
    # reference result: the whole stream on this GPU alone
    ms = n_bursts + 16
    ds = torch.zeros(ms * 16, dtype=torch.uint8, device=dev)
    dt = torch.zeros(ms * 288, dtype=torch.uint8, device=dev)
    dp = torch.zeros(ms * 9, dtype=torch.int32, device=dev)
    ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(full.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()),
                                   C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), ms)
    assert ns == n_bursts - 1, (ns, g.err())
    # the sharded run: only rank 0 hands its copy in
    handle = None
    src = full[:nbits] if rank == 0 else None
    if peer:
        # no scatter: rank 0 puts the stream into exportable memory, the others map it and read it over NVLink
        hb = torch.zeros(64, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf = T.DevBuffer(g, nbits + 64)
            src = buf.tensor(dev)[:nbits]
            src.copy_(full[:nbits])
            hb.copy_(torch.frombuffer(bytearray(buf.export()), dtype=torch.uint8))
        dist.broadcast(hb, 0)
        handle = bytes(hb.cpu().numpy().tobytes())
    k0, k1, a0, s_slots, s_t1, s_pk, summaries = T.sharded_decode(
        g, dist, rank, world, src, nbits, dev, want_type1=True, peer_handle=handle)
    n = k1 - k0
    ok = (torch.equal(s_slots[:n * 16], ds[k0 * 16:k1 * 16]) and torch.equal(s_t1[:n * 288], dt[k0 * 288:k1 * 288])
          and torch.equal(s_pk[:n * 9], dp[k0 * 9:k1 * 9]))
    codes = s_slots[:n * 16].view(torch.int32).view(-1, 4)[:, 1].unique().numel()
    tot = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    q.put((rank, bool(ok), n, int(tot[0]), ns, codes, all(s.first_unlock == 0xffffffff for s in summaries)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,peer", [(2, False), (2, True), (4, False), (8, False), (8, True)])
def test_one_stream_scattered_over_gpus(world, peer):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n_bursts = 400_000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bursts, q, peer)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok, n, tot, ns, codes, locked in res:
        assert ok, f"rank {rank}: shard differs from the single-GPU decode"
        assert tot == ns                       # the shards together are every slot of the stream
        assert codes > 1000                    # random cells: the carried cell state really mattered
        assert locked
