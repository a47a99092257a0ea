#!/usr/bin/env python
"""bench.py - TETRA bursts/s decoded (type-5 -> type-1 bits) on N B200s, and the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA path through the C ABI)
  python bench.py --impl reference ...                          the reference's own CPU code (oracle/_ref)
  python bench.py --workload config5 [--peer] ...               BASELINE config 5: one stream on rank 0 decoded by all ranks
  N > 1 is launched by the driver under torch.distributed.run, one rank per GPU.

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1], 10^6 synthetic SCH/F
bursts (RCPC 2/3, K=5 Viterbi) as one continuous downlink stream (two leading SYNC bursts give lock
and the cell code, an SB every 64th burst after that), BER 1e-2 on the payload bits.

  value         bursts/s with the stream already resident in HBM (tb200_rx_stream_dev); N > 1: every rank
                decodes its own stream (weak scaling, no data-path collective)
  e2e           the same through tb200_rx_stream_host: pinned HOST buffers in and out, H2D and D2H
                copies inside the timed region
  roofline      the dominant kernel against the measured HBM copy peak (+ integer-issue view), the
                training-sequence search kernel and the stand-alone descramble + de-interleave stage
  cpu_baseline  the reference's lower MAC compiled in place (oracle/_ref), all host cores, bounded sample
  other_configs device-resident bursts/s on the shapes of BASELINE configs 3 and 4 (N = 1 only)
  front_ends    the same stream bit-packed and as float32 symbols (N = 1 only)
  gsmtap_framing  GSMTAP frames of the step's CRC-good blocks built on the device (N = 1 only)
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "tetra_bursts_per_sec_decoded_bit_exact"
UNIT = "bursts/s"
WORKLOAD = "config2: 1e6 synthetic SCH/F bursts (RCPC 2/3, K=5 Viterbi), BER 1e-2, 64-byte reads"
N_BURSTS = 1_000_000
BYTES_PER_BURST_IN = 510           # 1 bit per byte input
BYTES_PER_BURST_OUT = 282          # type-1 bits of an SCH/F burst, 1 bit per byte
ACS_PER_BURST = 4672               # 16 states x 292 trellis steps
SEED = 0x7E7A0002


def gen_cfg(T, seed=SEED):
    return T.GenCfg(seed=seed, sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655,
                    random_cell=0, lead_in_bits=0)


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU reference arm

def _cpu_worker(args):
    """one process = one reference receiver (it is single-threaded with global state)"""
    seed, n_bursts, use_ref = args
    import tetra_testlib as T
    orc = T.Oracle()
    bits = orc.gen_stream(gen_cfg(T, seed), 0, n_bursts)
    if use_ref:
        rx = T.Ref()
    else:
        rx = orc
    rx.reset()
    rx.set_recording(False)
    t0 = time.perf_counter()
    rx.feed(bits, 64)
    dt = time.perf_counter() - t0
    return n_bursts, dt


def cpu_reference_rate(n_bursts_per_core, cores, one_core=False):
    """bursts/s of the reference's CPU path on `cores` processes, each over its own self-contained stream"""
    import tetra_testlib as T
    T.ensure_oracle_built()
    use_ref = T.have_ref()
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(SEED + 1000 + i, n_bursts_per_core, use_ref) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    one = None
    if one_core:
        with ctx.Pool(1) as pool:                          # the same sample on ONE core, nothing else running
            nb, dt1 = pool.map(_cpu_worker, [(SEED + 999, n_bursts_per_core, use_ref)])[0]
        one = nb / dt1
    return {"value": total / slowest, "unit": UNIT, "cores": cores, "one_core": one,
            "kind": "reference" if use_ref else "port",
            "sample": f"{cores} processes x {n_bursts_per_core} bursts of the bench workload (own seed each), "
                      f"64-byte reads, stdout silenced, {'oracle/_ref = reference lower MAC compiled in place + restated osmo_conv_decode (libosmocore absent)' if use_ref else 'oracle port'}; "
                      f"slowest process {slowest:.2f} s, pool wall {wall:.2f} s"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 20000
    for _ in range(args.warmup):
        cpu_reference_rate(2000, cores)
    t0 = time.perf_counter()
    rates = [cpu_reference_rate(per_core, cores) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = sum(r["value"] for r in rates) / len(rates)
    base = dict(rates[-1]); base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_bursts_per_step": per_core * cores},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------ our arm

def rank_seed(rank):
    """every rank decodes its own self-contained stream: shards are independent"""
    return SEED + rank


def reduce_max(dist, values, device):
    """max over ranks of a list of floats (device timings are reported as the slowest rank's)"""
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import tetra_testlib as T
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    G.load_package().load_library()          # fails loudly if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    g = T.B200(device=local_rank)
    n = N_BURSTS
    nbits = 510 * n
    cfg = gen_cfg(T, rank_seed(rank))
    d_bits = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
    rc = g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 0)
    assert rc == 0, g.err()
    ms = n + 16
    d_slots = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
    d_t1 = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
    g.set_options(chunk_bits=64, viterbi=args.viterbi, output=T.OUT_UNPACKED, pipeline_slots=0, profile=1)

    def step_dev():
        ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()),
                                       C.c_void_p(d_t1.data_ptr()), None, ms)
        assert ns == n - 1, (ns, g.err())
        return ns

    h_bits_p = g.lib.tb200_host_alloc(nbits)
    h_slots_p = g.lib.tb200_host_alloc(ms * 16)
    h_t1_p = g.lib.tb200_host_alloc(ms * 288)
    h_pk_p = g.lib.tb200_host_alloc(ms * 36)
    assert h_bits_p and h_slots_p and h_t1_p and h_pk_p
    h_bits = np.ctypeslib.as_array(C.cast(h_bits_p, C.POINTER(C.c_uint8)), shape=(nbits,))
    h_bits[:] = d_bits[:nbits].cpu().numpy()

    def step_host(packed=False):
        ns = g.lib.tb200_rx_stream_host(g.h, h_bits_p, nbits, 3, h_slots_p, None if packed else h_t1_p,
                                        h_pk_p if packed else None, ms)
        assert ns == n - 1, (ns, g.err())

    # ---- warm-up of both legs, then the clock sampler runs across both timed regions
    for _ in range(max(args.warmup, 3)):
        step_dev()
    g.set_options(profile=0)
    for _ in range(3):
        step_host()
    g.set_options(profile=1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- device-resident: value + per-kernel device times (CUDA events on the launching stream)
    barrier()
    t0 = time.perf_counter()
    tim = {"total": 0.0, "classify": 0.0, "scan": 0.0, "decode": 0.0, "search": 0.0}
    for _ in range(args.steps):
        step_dev()
        t = g.timing()
        tim["total"] += t.total_ms; tim["classify"] += t.classify_ms; tim["scan"] += t.scan_ms; tim["decode"] += t.decode_ms; tim["search"] += t.search_ms
    barrier()
    wall = time.perf_counter() - t0
    launches = g.stats().kernel_launches * args.steps     # stats restart with every TB200_FRESH call

    # ---- end to end: pinned host buffers, H2D + D2H inside the timed region
    g.set_options(profile=0, output=T.OUT_UNPACKED)
    wall_e2e = wall_e2e_packed = float("nan")
    if args.no_e2e:
        step_host()
    else:
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        wall_e2e = time.perf_counter() - t1
    hs = np.ctypeslib.as_array(C.cast(h_slots_p, C.POINTER(C.c_uint8)), shape=((n - 1) * 16,))
    ht = np.ctypeslib.as_array(C.cast(h_t1_p, C.POINTER(C.c_uint8)), shape=((n - 1) * 288,))
    same = bool(np.array_equal(hs, d_slots[:(n - 1) * 16].cpu().numpy())) and \
        bool(np.array_equal(ht[:288 * 4096], d_t1[:288 * 4096].cpu().numpy()))
    if not args.no_e2e:
        g.set_options(output=T.OUT_PACKED)
        for _ in range(2):
            step_host(packed=True)
        barrier()
        t2 = time.perf_counter()
        for _ in range(args.steps):
            step_host(packed=True)
        barrier()
        wall_e2e_packed = time.perf_counter() - t2
    # ---- input front ends: the same stream bit-packed (64 B per burst) and as float32 symbols
    front = None
    if not args.no_e2e and world == 1:
        front = {}
        nb8 = (nbits + 7) // 8
        pk_host = np.packbits(h_bits, bitorder="little")
        h_pk_in_p = g.lib.tb200_host_alloc(nb8 + 64)
        h_pk_in = np.ctypeslib.as_array(C.cast(h_pk_in_p, C.POINTER(C.c_uint8)), shape=(nb8,))
        h_pk_in[:] = pk_host
        d_pk_in = torch.from_numpy(pk_host).cuda()
        d_pk_in = torch.cat([d_pk_in, torch.zeros(64, dtype=torch.uint8, device="cuda")])
        code = d_bits[:nbits].view(-1, 2).to(torch.int64)
        code = code[:, 0] * 2 + code[:, 1]
        d_sym = (torch.tensor([1.0, 3.0, -1.0, -3.0], device="cuda")[code] +
                 (torch.rand(code.numel(), device="cuda") - 0.5) * 1.9).to(torch.float32).contiguous()
        del code

        def timed(fn, k):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(k):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t) / k

        def dev_step(buf):
            ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(buf.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()),
                                           C.c_void_p(d_t1.data_ptr()), None, ms)
            assert ns == n - 1, (ns, g.err())

        def host_step_packed_in(packed_out):
            ns = g.lib.tb200_rx_stream_host(g.h, h_pk_in_p, nbits, 3, h_slots_p, None if packed_out else h_t1_p,
                                            h_pk_p if packed_out else None, ms)
            assert ns == n - 1, (ns, g.err())
        k = max(10, args.steps // 2)
        g.set_options(profile=0, input=T.IN_PACKED, output=T.OUT_UNPACKED)
        t_dev_pk = timed(lambda: dev_step(d_pk_in), k)
        t_host_pk = timed(lambda: host_step_packed_in(False), k)
        same_pk = bool(np.array_equal(np.ctypeslib.as_array(C.cast(h_slots_p, C.POINTER(C.c_uint8)), shape=((n - 1) * 16,)),
                                      d_slots[:(n - 1) * 16].cpu().numpy()))
        g.set_options(output=T.OUT_PACKED)
        t_host_pk_pk = timed(lambda: host_step_packed_in(True), k)
        g.set_options(input=T.IN_F32SYM, output=T.OUT_UNPACKED)
        t_dev_sym = timed(lambda: dev_step(d_sym), k)
        g.set_options(input=T.IN_BYTES, output=T.OUT_UNPACKED)
        front = {"packed_input": {"device_resident": {"value": (n - 1) / t_dev_pk, "unit": UNIT, "ms_per_step": t_dev_pk * 1e3},
                                  "e2e_host_buffers": {"value": (n - 1) / t_host_pk, "unit": UNIT, "ms_per_step": t_host_pk * 1e3,
                                                       "h2d_bytes_per_step": nb8, "d2h_bytes_per_step": (n - 1) * (16 + 288),
                                                       "matches_device_path": same_pk},
                                  "e2e_host_buffers_packed_output": {"value": (n - 1) / t_host_pk_pk, "unit": UNIT,
                                                                     "ms_per_step": t_host_pk_pk * 1e3, "h2d_bytes_per_step": nb8,
                                                                     "d2h_bytes_per_step": (n - 1) * (16 + 36)},
                                  "format": "8 stream bits per byte (TB200_IN_PACKED), 64 B per burst"},
                 "symbol_input": {"device_resident": {"value": (n - 1) / t_dev_sym, "unit": UNIT, "ms_per_step": t_dev_sym * 1e3},
                                  "format": "float32 per symbol (TB200_IN_F32SYM), sliced on the device like float_to_bits.c, 1020 B per burst"}}
        g.lib.tb200_host_free(h_pk_in_p)
        del d_pk_in, d_sym
    clocks = sampler.stop() if rank == 0 else None

    # ---- the fused descramble + de-interleave stage on its own (north star: >= 70 % of the HBM roofline)
    g.set_options(profile=1)
    nblk = 4_000_000
    d5 = torch.randint(0, 2, (nblk * 432,), dtype=torch.uint8, device="cuda")
    d3 = torch.empty_like(d5)
    dcodes = torch.randint(0, 2 ** 31 - 1, (nblk,), dtype=torch.int32, device="cuda")
    stage_ms = []
    for _ in range(6):
        rc = g.lib.tb200_descramble_deinterleave(g.h, C.c_void_p(d5.data_ptr()), C.c_void_p(d3.data_ptr()),
                                                 C.c_void_p(dcodes.data_ptr()), nblk, 432, 103, 1)
        assert rc == 0, g.err()
        stage_ms.append(g.timing().leaf_ms)
    stage_ms = min(stage_ms[1:])
    del d5, d3

    # ---- GSMTAP framing of the decoded blocks: slot records + packed type-1 words -> frames
    gsmtap = None
    if not args.no_e2e and world == 1 and hasattr(g.lib, "tb200_gsmtap_pack"):
        d_pk = torch.zeros(ms * 9, dtype=torch.int32, device="cuda")
        g.set_options(profile=0, input=T.IN_BYTES, output=T.OUT_PACKED)
        ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()), None,
                                       C.c_void_p(d_pk.data_ptr()), ms)
        assert ns == n - 1, (ns, g.err())
        nf = C.c_uint64(0)
        need = g.lib.tb200_gsmtap_pack(g.h, C.c_void_p(d_slots.data_ptr()), None, ns, None, 0, None, C.byref(nf), 1)
        assert need > 0, g.err()
        d_fr = torch.empty(need, dtype=torch.uint8, device="cuda")
        g.set_options(profile=1)
        gt_ms = []
        for _ in range(6):
            rc = g.lib.tb200_gsmtap_pack(g.h, C.c_void_p(d_slots.data_ptr()), C.c_void_p(d_pk.data_ptr()), ns,
                                         C.c_void_p(d_fr.data_ptr()), need, None, None, 1)
            assert rc == need, g.err()
            gt_ms.append(g.timing().leaf_ms)
        gt_ms = min(gt_ms[1:])
        gt_bytes = ns * (16 + 16 + 36) + need          # slot record read twice (sizes pass + emit pass), packed words, frames
        gsmtap = {"kernels": "k_gsmtap_sizes + k_gsmtap_scan + k_gsmtap_emit", "slots": int(ns), "frames": int(nf.value),
                  "frame_bytes": int(need), "ms": gt_ms, "slots_per_s": ns / (gt_ms * 1e-3), "bound": "hbm",
                  "achieved": gt_bytes / (gt_ms * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes": int(gt_bytes)}
        del d_pk, d_fr
    g.set_options(profile=1, input=T.IN_BYTES, output=T.OUT_UNPACKED)

    wall, wall_e2e, wall_e2e_packed, t_total, t_cls, t_scan, t_dec, t_search = reduce_max(
        dist, [wall, wall_e2e, wall_e2e_packed, tim["total"], tim["classify"], tim["scan"], tim["decode"], tim["search"]], "cuda")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    bursts = (n - 1) * args.steps * world
    value = bursts / wall
    e2e = bursts / wall_e2e
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    per_launch_dec_ms = t_dec / args.steps
    per_launch_cls_ms = t_cls / args.steps
    per_launch_search_ms = t_search / args.steps
    alg_bytes = (BYTES_PER_BURST_IN + BYTES_PER_BURST_OUT) * (n - 1)
    dec_gbs = alg_bytes / (per_launch_dec_ms * 1e-3) / 1e9
    search_gbs = BYTES_PER_BURST_IN * (n - 1) / (per_launch_search_ms * 1e-3) / 1e9
    stage_gbs = 2 * 432 * nblk / (stage_ms * 1e-3) / 1e9
    int_peak = g.lib.tb200_measure_int_peak(g.h)
    acs_rate = ACS_PER_BURST * (n - 1) / (per_launch_dec_ms * 1e-3)
    dec_name = "k_decode_warp" if args.viterbi == 0 else "k_decode_lane"
    # per-launch constants read from the committed ncu capture of this same workload (profiles/kernel_constants.json)
    kc = {}
    try:
        kc = json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))
    except Exception:
        pass
    kdec, kcls = kc.get(dec_name, {}), kc.get("k_classify_tile", {})
    inst_per_acs = (kdec["thread_inst_per_launch"] / (ACS_PER_BURST * (n - 1))) if "thread_inst_per_launch" in kdec else None
    roofline = {"kernel": dec_name, "bound": "hbm",
                "achieved": dec_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dec_gbs / hbm_peak,
                "traffic": kdec.get("dram_bytes_per_launch"),
                "peak_source": peak_src, "ms_per_launch": per_launch_dec_ms,
                "algorithmic_bytes_per_burst": BYTES_PER_BURST_IN + BYTES_PER_BURST_OUT,
                "note": "the decode kernel is integer-issue bound (16-state add-compare-select), not HBM bound: int_alu is its real roofline",
                "int_alu": {"acs_per_s": acs_rate, "thread_inst_per_acs": inst_per_acs,
                            "int_inst_per_s": (acs_rate * inst_per_acs) if inst_per_acs else None,
                            "measured_int_peak_inst_per_s": int_peak,
                            "frac": (acs_rate * inst_per_acs / int_peak) if (inst_per_acs and int_peak) else None,
                            "how": "thread instructions the kernel executes per launch (ncu smsp__inst_executed x 32, profiles/kernel_constants.json) / kernel time, "
                                   "against a register-only add+min kernel that keeps both integer pipes busy (tb200_measure_int_peak); "
                                   "4672 add-compare-select per SCH/F burst, each a packed add + VIADDMNMX.U16x2 shared by two trellises"},
                "sync_search": {"kernel": "k_classify_tile" if args.viterbi else "k_classify", "bound": "hbm", "achieved": search_gbs,
                                "peak": hbm_peak, "unit": "GB/s", "frac": search_gbs / hbm_peak, "ms_per_launch": per_launch_search_ms,
                                "algorithmic_bytes_per_burst": BYTES_PER_BURST_IN, "traffic": kcls.get("dram_bytes_per_launch"),
                                "with_sb1_pass_ms": per_launch_cls_ms},
                "descramble_deinterleave_stage": {"kernel": "k_stage_tma", "bound": "hbm", "achieved": stage_gbs,
                                                  "peak": hbm_peak, "unit": "GB/s", "frac": stage_gbs / hbm_peak, "ms_per_launch": stage_ms,
                                                  "algorithmic_bytes_per_block": 864, "blocks": nblk},
                "step_share": {"classify": t_cls / t_total, "scan": t_scan / t_total, "decode": t_dec / t_total},
                "constants_from": kc.get("source")}
    cores = os.cpu_count() or 1
    cpu = cpu_reference_rate(40000, cores, one_core=True) if (world == 1 and not args.no_cpu) else None
    others = other_configs(g, T, torch, C) if (world == 1 and not args.no_e2e) else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bursts_per_gpu_per_step": n,
                       "viterbi": "warp-shuffle, one warp per burst" if args.viterbi == 0 else "lane: two packed trellises per thread",
                       "l2": "inputs larger than L2 (510 MB stream per step)", "parallelism": f"independent streams x{world}",
                       "output": "slot records + unpacked type-1 bits (1 bit/byte)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": nbits * world,
                    "d2h_bytes_per_step": (n - 1) * (16 + 288) * world, "ms_per_step": wall_e2e / args.steps * 1e3,
                    "matches_device_path": same,
                    "packed_output": {"value": bursts / wall_e2e_packed, "d2h_bytes_per_step": (n - 1) * (16 + 36) * world,
                                      "ms_per_step": wall_e2e_packed / args.steps * 1e3}},
            "gpu_launches": int(launches), "device_ms_per_step": t_total / args.steps,
            "roofline": roofline, "clocks": clocks}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if others is not None:
        line["other_configs"] = others
    if front is not None:
        line["front_ends"] = front
    if gsmtap is not None:
        gsmtap["peak"] = hbm_peak
        gsmtap["frac"] = gsmtap["achieved"] / hbm_peak
        line["gsmtap_framing"] = gsmtap
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def other_configs(g, T, torch, C):
    """device-resident bursts/s on the shapes of BASELINE configs 3 and 4 (not the headline: context only)"""
    out = {}
    shapes = {"config3: mixed SB/NDB(1 and 2 channel) bursts, 333-bit lead-in, lock FSM, random cells, BER 1e-2":
                  dict(sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333),
              "config4: every other burst a SYNC burst announcing a random cell (per-burst scrambling codes), AACH RM(30,14), BER 1e-2":
                  dict(sb_period=2, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)}
    n = 2_000_000
    for name, kw in shapes.items():
        cfg = T.GenCfg(seed=0x7E7A0003, **kw)
        nbits = 510 * n + kw["lead_in_bits"]
        d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
        assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0, g.err()
        ms = n + 16
        ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
        dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
        g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=0)
        def step():
            ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()),
                                           C.c_void_p(dt.data_ptr()), None, ms)
            assert ns > 0.99 * n, (ns, g.err())
            return ns
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 20
        for _ in range(k):
            ns = step()
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        st = g.stats()
        flags = ds[:ns * 16].view(-1, 16)[:, 15]
        kinds = torch.bincount((flags & 3).to(torch.int64), minlength=4).cpu().tolist()
        out[name] = {"value": ns * k / dt_s, "unit": UNIT, "bursts_per_step": n, "slots_decoded": int(ns),
                     "kinds": {"dropped": kinds[0], "sync": kinds[1], "ndb_schf": kinds[2], "ndb_two_blocks": kinds[3]},
                     "lock_losses": int(st.lock_losses), "ms_per_step": dt_s / k * 1e3}
        del d, ds, dt
    return out


def run_config5(args, rank, world, local_rank):
    """BASELINE config 5: ONE stream of --total-bursts bursts (config-4 shape) held by rank 0, scattered over the
    ranks with NCCL send/recv, pass 1 per shard, all-gather of the 32-byte cell-state summaries, pass 2.
    Strong scaling: the total is fixed.  The timed region of `value` contains the scatter; `decode_only`
    excludes it (shards already resident on their GPUs)."""
    import torch
    import tetra_testlib as T
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    G.load_package().load_library()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    g = T.B200(device=local_rank)
    n = args.total_bursts
    cfg = T.GenCfg(seed=0x7E7A0005, sb_period=2, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333)
    nbits = 510 * n + 333
    full = None
    handle = None
    if rank == 0:
        buf = T.DevBuffer(g, nbits + 64)            # exportable memory: the peer mode maps it into the other ranks
        full = buf.tensor(dev)
        assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(full.data_ptr()), 1) == 0, g.err()
        full = full[:nbits]
    if args.peer and world > 1:
        hb = torch.zeros(64, dtype=torch.uint8, device=dev)
        if rank == 0:
            hb.copy_(torch.frombuffer(bytearray(buf.export()), dtype=torch.uint8))
        dist.broadcast(hb, 0)
        handle = bytes(hb.cpu().numpy().tobytes())
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, pipeline_slots=0, output=T.OUT_PACKED)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    tot_s = dec_s = 0.0
    slots_total = 0
    for it in range(args.warmup + args.steps):
        if it == args.warmup and rank == 0:
            sampler.start()
        tm = {}
        barrier()
        t0 = time.perf_counter()
        k0, k1, a0, d_slots, _, d_pk, summ = T.sharded_decode(g, dist, rank, world, full, nbits, dev, timers=tm, peer_handle=handle)
        barrier()
        t1 = time.perf_counter()
        if it >= args.warmup:
            tot_s += t1 - t0
            dec_s += (t1 - t0) - (tm["t_scatter1"] - tm["t_scatter0"])
            slots_total = sum(s.n_slots for s in summ)
        del d_slots, d_pk
    clocks = sampler.stop() if rank == 0 else None
    tot_s, dec_s = reduce_max(dist, [tot_s, dec_s], "cuda")
    if rank == 0:
        line = {"metric": METRIC, "value": slots_total * args.steps / tot_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"config5: one stream of {n} bursts (config-4 shape) on rank 0, " +
                                       ("shards read in place over NVLink by the search kernels (peer-mapped memory)" if handle else
                                        "NCCL scatter of contiguous shards") + " + all-gather of 32-byte cell-state summaries", "total_bursts": n,
                           "l2": "inputs larger than L2", "parallelism": f"one stream sharded x{world}", "output": "slot records + packed type-1 bits, rank-local"},
                "decode_only": {"value": slots_total * args.steps / dec_s, "unit": UNIT, "ms_per_step": dec_s / args.steps * 1e3,
                                "note": "scatter excluded (wall-clock between device synchronisations, max over ranks)"},
                "scatter_bytes_per_step": int((nbits) * (world - 1) / world), "clocks": clocks}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--viterbi", type=int, default=1, help="0 warp-shuffle (one warp per burst), 1 lane (two trellises per thread)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--no-cpu", action="store_true", help="profiling runs: skip the CPU baseline leg")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2: headline (independent streams per GPU); config5: one stream scattered over the GPUs with NCCL")
    ap.add_argument("--total-bursts", type=int, default=100_000_000, help="config5: bursts in the one stream")
    ap.add_argument("--peer", action="store_true", help="config5: no scatter, every rank's search kernel reads its shard from rank 0's memory over NVLink")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "config5":
        if args.steps == 100:
            args.steps = 5
        run_config5(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
