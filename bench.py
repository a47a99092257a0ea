#!/usr/bin/env python
"""bench.py - TETRA bursts/s decoded (type-5 -> type-1 bits) on N B200s, and the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA path through the C ABI)
  python bench.py --impl reference ...                          the reference's own CPU code (oracle/_ref)
  N > 1 is launched by the driver under torch.distributed.run, one rank per GPU.

A "step" is one pass of the hot path over one batch.  The headline workload is the largest single-GPU configuration of
BASELINE.json: config 4, 10^8 bursts as ONE continuous downlink stream (51 GB at the reference's one-byte-per-bit ABI,
generated on the device): every other burst a SYNC burst that announces a random cell (MCC / MNC / colour code), so the
scrambling code of every block is learned from the stream itself (tetra_lower_mac.c:291-302), AACH = RM(30,14) code
words, 25 % two-channel normal bursts, 333 random bits in front, BER 1e-2 on the payload bits, 64-byte reads.

  value         bursts/s with the stream already resident in HBM (tb200_rx_stream_dev); N > 1: every rank decodes
                its own stream (weak scaling, no data-path collective)
  e2e           the same through tb200_rx_stream_host: pinned HOST buffers in and out, H2D and D2H copies inside
                the timed region
  parity        windows of the TIMED output replayed through the reference's own C (oracle/_ref) + whole-run digest
                of the device path against the host path; a mismatch fails the run
  roofline      the dominant kernel (decode pass, integer-issue bound) + the search kernel (HBM bound) + the stand-alone
                descramble / de-interleave stage
  configs       config 2 (10^6 SCH/F bursts) and config 3 (10^7 mixed bursts) as named sub-records with their own rooflines
  config5       BASELINE config 5: rank 0's stream sharded over all N ranks by the C driver (tb200_dist_rx_stream: on-device
                packing, NCCL scatter or peer reads, one all-gather), strong scaling, parity-checked
  cpu_baseline  the reference's lower MAC compiled in place (oracle/_ref), all host cores, bounded sample (N = 1)
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
N_BURSTS = 100_000_000
WORKLOAD = ("config4: 1e8 bursts, one continuous stream, every other burst a SYNC burst announcing a random cell "
            "(per-burst scrambling codes learned from the stream), AACH RM(30,14), 25% two-channel bursts, BER 1e-2, 64-byte reads")
BYTES_PER_BURST_IN = 510           # 1 bit per byte input
SEED = 0x7E7A0004
# add-compare-selects per block (16 states x (type-2 bits + 4 flush steps))
ACS_SB1, ACS_HALF, ACS_SCHF = 16 * 84, 16 * 148, 16 * 292
# type-1 bytes per slot kind at one bit per byte: SB 60+14+124, NDB/SCH-F 14+268, two-channel 14+124+124
T1_BYTES = {1: 198, 2: 282, 3: 262, 0: 0}

SHAPES = {
    "config2": dict(sb_period=64, lead_sb=2, ndb2_per_256=0, ber_per_65536=655, random_cell=0, lead_in_bits=0),
    "config3": dict(sb_period=18, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333),
    "config4": dict(sb_period=2, lead_sb=2, ndb2_per_256=64, ber_per_65536=655, random_cell=1, lead_in_bits=333),
}
NAMES = {
    "config2": "config2: 1e6 synthetic SCH/F bursts (RCPC 2/3, K=5 Viterbi), BER 1e-2, 64-byte reads",
    "config3": "config3: 1e7 mixed SB / NDB (one and two channel) bursts, 333-bit lead-in, lock FSM, random cells, BER 1e-2",
    "config4": WORKLOAD,
}


def gen_cfg(T, seed=SEED, shape="config4"):
    return T.GenCfg(seed=seed, **SHAPES[shape])


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
    seed, n_bursts, use_ref, shape = args
    import tetra_testlib as T
    orc = T.Oracle()
    bits = orc.gen_stream(gen_cfg(T, seed, shape), 0, n_bursts)
    rx = T.Ref() if use_ref else orc
    rx.reset()
    rx.set_recording(False)
    t0 = time.perf_counter()
    rx.feed(bits, 64)
    dt = time.perf_counter() - t0
    return n_bursts, dt


def cpu_reference_rate(n_bursts_per_core, cores, one_core=False, shape="config4"):
    """bursts/s of the reference's CPU path on `cores` processes, each over its own self-contained stream"""
    import tetra_testlib as T
    T.ensure_oracle_built()
    use_ref = T.have_ref()
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(SEED + 1000 + i, n_bursts_per_core, use_ref, shape) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    one = None
    if one_core:
        with ctx.Pool(1) as pool:                          # the same sample on ONE core, nothing else running
            nb, dt1 = pool.map(_cpu_worker, [(SEED + 999, n_bursts_per_core, use_ref, shape)])[0]
        one = nb / dt1
    return {"value": total / slowest, "unit": UNIT, "cores": cores, "one_core": one,
            "kind": "reference" if use_ref else "port",
            "sample": f"{cores} processes x {n_bursts_per_core} bursts of the bench workload's shape ({shape}, own seed each), "
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


def bind_to_gpu_numa_node(local_rank):
    """run this rank (and first-touch its pinned buffers) on the cores next to its GPU; no-op where the box shows one node"""
    try:
        bus = subprocess.check_output(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip()
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        if node < 0:
            return {"numa_node": node, "bound": False}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return {"numa_node": node, "bound": True, "cpus": cpus}
    except Exception as e:                         # best effort: an odd sysfs layout must not fail the bench
        return {"numa_node": None, "bound": False, "why": str(e)[:80]}


class Peaks:
    def __init__(self):
        self.d = {}
        try:
            self.d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm = float(self.d.get("hbm_gbs", 6650.0))
        self.src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in self.d else "fallback 6650 GB/s (B200_PROFILING.md)"


def kernel_constants():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))
    except Exception:
        return {}


TIM_KEYS = ("total", "classify", "scan", "decode", "search", "prepare", "trellis")


def add_timing(tim, t):
    tim["total"] += t.total_ms; tim["classify"] += t.classify_ms; tim["scan"] += t.scan_ms; tim["decode"] += t.decode_ms
    tim["search"] += t.search_ms; tim["prepare"] += t.prepare_ms; tim["trellis"] += t.trellis_ms


def serial_kernel_times(g, step, steps):
    """per-kernel CUDA-event times of `steps` steps with options.serial_passes = 1, scaled to one step"""
    g.set_options(serial_passes=1, profile=1)
    step()
    tim = dict.fromkeys(TIM_KEYS, 0.0)
    for _ in range(steps):
        step()
        t = g.timing()
        add_timing(tim, t)
    g.set_options(serial_passes=0)
    return {k: v / steps for k, v in tim.items()}


def kinds_of(torch, d_slots, ns):
    flags = d_slots[:ns * 16].view(-1, 16)[:, 15]
    return torch.bincount((flags & 3).to(torch.int64), minlength=4).cpu().tolist()


def rooflines(kinds, ns, tim, steps, int_peak, peaks, kc, shape):
    """decode pass against the integer-issue peak, search kernel against HBM, from CUDA-event times (sums over the pieces of
    a step; tim holds totals over `steps` steps, taken with serial_passes = 1) and the kind mix of the decoded stream"""
    dec_ms, search_ms, cls_ms, scan_ms, total_ms = (tim[k] / steps for k in ("decode", "search", "classify", "scan", "total"))
    prep_ms, trel_ms = tim.get("prepare", 0.0) / steps, tim.get("trellis", 0.0) / steps
    split = trel_ms > 0.05 * dec_ms                     # the split decode pass (k_lane_prepare | k_lane_trellis), the default form
    kname = "k_lane_trellis" if split else "k_decode_lane"
    k_ms = trel_ms if split else dec_ms                 # the dominant kernel's own time
    acs_decode = kinds[1] * ACS_HALF + kinds[2] * ACS_SCHF + kinds[3] * 2 * ACS_HALF          # SB2, SCH/F, BLK1 + BLK2
    acs_sb1 = kinds[1] * ACS_SB1
    out_bytes = sum(kinds[k] * T1_BYTES[k] for k in range(4))
    acs_rate = acs_decode / (k_ms * 1e-3)
    kd = kc.get(kname, {}).get(shape, {})
    inst_per_acs = kd.get("thread_inst_per_acs")
    dec = {"kernel": kname, "bound": "int_issue",
           "achieved": acs_rate / 1e9, "peak": int_peak / 1e9, "unit": "G/s (add-compare-selects against integer thread-instructions)",
           "frac": acs_rate / int_peak if int_peak else None,
           "traffic": kd.get("dram_bytes_per_slot") and kd["dram_bytes_per_slot"] * ns,
           "ms_per_step": k_ms, "acs_per_step": acs_decode,
           "how": "algorithmic ACS of the blocks the pass decodes (16 states x (type-2 bits + 4) per block) / summed CUDA-event time of the kernel "
                  "that runs them (k_lane_trellis: ACS loop + trace back + CRC / type-1 output); the packed form needs one thread instruction per "
                  "ACS (add + VIADDMNMX.U16x2 per state for two trellises), so frac is the share of the measured integer issue peak "
                  "(tb200_measure_int_peak: register-only add+min on both pipes) spent on ACS proper",
           "thread_inst_per_acs": inst_per_acs,
           "issue_utilisation": (acs_rate * inst_per_acs / int_peak) if (inst_per_acs and int_peak) else None,
           "decode_pass": {"ms_per_step": dec_ms, "k_lane_prepare_ms": prep_ms if split else None, "k_lane_trellis_ms": trel_ms if split else None,
                           "frac": acs_decode / (dec_ms * 1e-3) / int_peak if int_peak else None,
                           "note": "the whole pass 2 (prepare: cell state + descramble + de-interleave, then the trellis kernel) on the same scale"},
           "hbm_view": {"achieved": (BYTES_PER_BURST_IN * ns + out_bytes) / (dec_ms * 1e-3) / 1e9, "peak": peaks.hbm, "unit": "GB/s",
                        "note": "algorithmic bytes of the whole chain (510 B in + type-1 bytes out per slot) over the decode time: the pass is not HBM bound"}}
    dec["hbm_view"]["frac"] = dec["hbm_view"]["achieved"] / peaks.hbm
    search_gbs = BYTES_PER_BURST_IN * ns / (search_ms * 1e-3) / 1e9
    ks = kc.get("k_classify_tile", {}).get(shape, {})
    own_gbs = (BYTES_PER_BURST_IN + 64 + 32) * ns / (search_ms * 1e-3) / 1e9
    search = {"kernel": "k_classify_tile", "bound": "hbm", "achieved": search_gbs, "peak": peaks.hbm, "unit": "GB/s",
              "frac": search_gbs / peaks.hbm, "ms_per_step": search_ms, "algorithmic_bytes_per_burst": BYTES_PER_BURST_IN,
              "with_its_outputs": {"bytes_per_burst": BYTES_PER_BURST_IN + 64 + 32, "achieved": own_gbs, "frac": own_gbs / peaks.hbm,
                                   "note": "510 B of stream read + the 64-byte packed slot record and the 32-byte slot state it has to write "
                                           "for the later passes; ncu measures 600 B of DRAM traffic per slot"},
              "traffic": ks.get("dram_bytes_per_slot") and ks["dram_bytes_per_slot"] * ns, "peak_source": peaks.src}
    return dec, search, {"search": search_ms / total_ms, "sb1": (cls_ms - search_ms) / total_ms, "scan": scan_ms / total_ms,
                         "decode": dec_ms / total_ms, "sb1_acs_per_step": acs_sb1}


def window_parity(T, torch, g, rx, d_bits, nbits, lead, d_slots, d_t1, ns, shape, n_windows, rng):
    """replay windows of the stream through the CPU checker (the reference's own C when oracle/_ref travelled) and
    compare, record for record, with the slots of the TIMED output.  A window starts at a SYNC burst: the checker gives
    that burst up for lock and is in step with the stream's receiver from the next CRC-good SB1 on."""
    import numpy as np
    sbp = SHAPES[shape]["sb_period"]
    win_bursts = max(200, 3 * sbp + 60)
    n_bursts = (nbits - lead) // 510
    checked = records = mismatches = 0
    for _ in range(n_windows):
        k0 = sbp * int(rng.integers(1, (n_bursts - win_bursts) // sbp - 1))
        a = lead + 510 * k0
        win = d_bits[a:a + 510 * win_bursts].cpu().numpy()
        rx.reset(); rx.feed(win, 64)
        want = rx.records()
        sel0, sel1 = k0, k0 + win_bursts - 2                 # slot i of the run is burst i + 1 (the first burst gives lock)
        sl = d_slots[sel0 * 16:sel1 * 16].cpu().numpy().view(T.SLOT_DTYPE).copy()
        t1 = d_t1[sel0 * 288:sel1 * 288].cpu().numpy().reshape(-1, 288)
        if not np.array_equal(sl["slot_bit"], ((lead + 510 * (np.arange(sel0, sel1) + 1)) & 0xffffffff).astype(np.uint32)):
            return {"checked": checked, "mismatches": 1, "note": "slot positions differ (lock lost inside the run?)"}
        sl["slot_bit"] = (510 * (np.arange(sel0, sel1) + 1 - k0)).astype(np.uint32)
        got = g.expand_records(sl, t1)
        ok = False
        for j in range(1, 2 * sbp + 40):                     # first burst from which the cold checker is in step
            lo, hi = 510 * j, 510 * (win_bursts - 2)
            w = want[(want["slot_bit"] >= lo) & (want["slot_bit"] < hi)]
            h = got[(got["slot_bit"] >= lo) & (got["slot_bit"] < hi)]
            ok, _ = T.records_equal(w, h)
            if ok:
                checked += win_bursts - 2 - j
                records += int(w.size)
                break
        if not ok:
            mismatches += 1
    return {"checked": checked, "records": records, "windows": n_windows, "mismatches": mismatches}


def run_shape(g, T, torch, shape, n, steps, warmup, seed, int_peak, peaks, kc, rx, rng):
    """device-resident decode of a smaller configuration as a named sub-record with its own rooflines and parity"""
    cfg = gen_cfg(T, seed, shape)
    lead = SHAPES[shape]["lead_in_bits"]
    nbits = 510 * n + lead
    d = torch.zeros(nbits + 64, dtype=torch.uint8, device="cuda")
    assert g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d.data_ptr()), 1) == 0, g.err()
    ms = n + 16
    ds = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
    dt = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=1, input=T.IN_BYTES)

    def step():
        ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d.data_ptr()), nbits, 3, C.c_void_p(ds.data_ptr()), C.c_void_p(dt.data_ptr()), None, ms)
        assert ns > 0.99 * n, (ns, g.err())
        return ns
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    tim = dict.fromkeys(TIM_KEYS, 0.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        ns = step()
        t = g.timing()
        add_timing(tim, t)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kinds = kinds_of(torch, ds, ns)
    dev_ms = tim["total"] / steps
    tim = serial_kernel_times(g, step, max(2, min(steps, 5)))
    dec, search, share = rooflines(kinds, ns, tim, 1, int_peak, peaks, kc, shape)
    st = g.stats()
    out = {"workload": NAMES[shape], "value": ns * steps / wall, "unit": UNIT, "bursts_per_step": n, "slots_decoded": int(ns), "steps": steps,
           "ms_per_step": wall / steps * 1e3, "device_ms_per_step": dev_ms, "device_ms_per_step_serial_passes": tim["total"],
           "kinds": {"dropped": kinds[0], "sync": kinds[1], "ndb_schf": kinds[2], "ndb_two_blocks": kinds[3]},
           "lock_losses": int(st.lock_losses), "crc_ok_blocks": int(st.crc_ok_blocks), "blocks": int(st.blocks),
           "roofline": dec, "sync_search": search, "step_share": share}
    if rx is not None and st.lock_losses == 0:
        out["parity"] = window_parity(T, torch, g, rx, d, nbits, lead, ds, dt, ns, shape, 3, rng)
    del d, ds, dt
    return out


def run_config5(g, T, torch, dist, rank, world, d_bits, nbits, rx, rng, steps, warmup):
    """BASELINE config 5: rank 0's 10^8-burst stream sharded over all ranks by the C driver.  Strong scaling: the total is
    fixed.  Parity: the digests of the ranks' runs add up to the digest of the single-GPU decode of the same stream, and
    every rank replays a window at its shard edge on the CPU checker."""
    import numpy as np
    dev = torch.device("cuda", torch.cuda.current_device())
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(T.Dist.get_id(g)), dtype=torch.uint8))
    if dist is not None:
        dist.broadcast(idt, 0)
    dd = T.Dist(g, rank, world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_PACKED, pipeline_slots=0, profile=0, input=T.IN_BYTES)
    mls = dd.max_local_slots(nbits)
    s_slots = torch.zeros(mls * 16, dtype=torch.uint8, device=dev)
    s_pk = torch.zeros(mls * 9, dtype=torch.int32, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # single-GPU reference of the same stream (rank 0, untimed): digest + the slots the edge windows are compared with
    one_digest = ns_one = 0
    if rank == 0:
        ms = nbits // 510 + 16
        o_slots = torch.zeros(ms * 16, dtype=torch.uint8, device=dev)
        o_pk = torch.zeros(ms * 9, dtype=torch.int32, device=dev)
        ns_one = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(o_slots.data_ptr()), None,
                                           C.c_void_p(o_pk.data_ptr()), ms)
        assert ns_one > 0, g.err()
        one_digest = g.slots_digest(o_slots.data_ptr(), o_pk.data_ptr(), 0, is_device=True, n=ns_one)
        del o_slots, o_pk
    modes = {"scatter_packed": T.DIST_SCATTER | T.DIST_PACK, "peer_packed": T.DIST_PEER | T.DIST_PACK}
    if world == 1:
        modes = {"scatter_packed": T.DIST_SCATTER | T.DIST_PACK}
    rec = {}
    for name, mode in modes.items():
        tot = 0.0
        phases = dict.fromkeys(("pack_ms", "lock_ms", "transfer_ms", "pass1_ms", "exchange_ms", "pass2_ms"), 0.0)
        for it in range(warmup + steps):
            barrier()
            t0 = time.perf_counter()
            n, runs = dd.rx_stream(d_bits.data_ptr() if rank == 0 else None, nbits, mode, s_slots.data_ptr(), None, s_pk.data_ptr(), mls)
            barrier()
            if it >= warmup:
                tot += time.perf_counter() - t0
                tm = dd.timing()
                for k in phases:
                    phases[k] += getattr(tm, k)
        dig = 0
        for gs, l, c in runs:
            dig = (dig + g.slots_digest(s_slots.data_ptr() + 16 * l, s_pk.data_ptr() + 36 * l, gs, is_device=True, n=c)) & (2 ** 64 - 1)
        # digests and slot counts of all ranks (split into 32-bit halves: NCCL sums int64 without surprises)
        v = torch.tensor([dig & 0xffffffff, dig >> 32, n, dd.timing().bytes_sent, dd.timing().segments], dtype=torch.int64, device=dev)
        allv = [torch.zeros_like(v) for _ in range(world)]
        if dist is not None:
            dist.all_gather(allv, v)
        else:
            allv = [v]
        allv = [[int(x) for x in t.cpu()] for t in allv]
        total_digest = sum(a[0] | (a[1] << 32) for a in allv) & (2 ** 64 - 1)
        n_total = sum(a[2] for a in allv)
        tot_max, *ph = reduce_max(dist, [tot] + [phases[k] for k in phases], dev)
        rec[name] = {"value": n_total * steps / tot_max, "unit": UNIT, "ms_per_step": tot_max / steps * 1e3, "slots": n_total,
                     "phases_ms_per_step_max_over_ranks": {k: p / steps for k, p in zip(phases, ph)},
                     "bytes_over_nvlink_per_step": allv[0][3], "segments": allv[0][4],
                     "digest_equals_single_gpu_decode": None, "_digest": total_digest, "_n": n_total}
    # CPU replay of a window that straddles this rank's shard edge (every rank can regenerate rank 0's stream: the device
    # generator has a CPU twin): the first slots of the shard are the ones whose cell state came over the all-gather
    edge_ok = 1
    if rx is not None and runs and allv[0][4] == 1:
        gs, l, c = runs[0]
        orc = T.Oracle()
        k0 = max(2, ((gs + 1 - 40) // 2) * 2) if rank else 2          # an even burst = a SYNC burst, 40 bursts before the edge
        nb = 200
        win = orc.gen_stream(gen_cfg(T, rank_seed(0), "config4"), k0, nb, lead_in=False)
        rx.reset(); rx.feed(win, 64)
        want = rx.records()
        i0 = max(gs, k0)                                            # global slot i is burst i + 1
        i1 = min(gs + c, k0 + nb - 2)
        sl = s_slots[(l + i0 - gs) * 16:(l + i1 - gs) * 16].cpu().numpy().view(T.SLOT_DTYPE).copy()
        pk = s_pk[(l + i0 - gs) * 9:(l + i1 - gs) * 9].cpu().numpy().view(np.uint32).reshape(-1, 9)
        sl["slot_bit"] = (510 * (np.arange(i0, i1) + 1 - k0)).astype(np.uint32)
        unp = ((pk[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(sl.size, 288).astype(np.uint8)
        got = g.expand_records(sl, unp)
        lo, hi = 510 * (i0 + 1 - k0), 510 * (nb - 2)
        w = want[(want["slot_bit"] >= lo) & (want["slot_bit"] < hi)]
        h = got[(got["slot_bit"] >= lo) & (got["slot_bit"] < hi)]
        edge_ok = int(T.records_equal(w, h)[0]) if rank else 1     # (rank 0's shard starts where the stream does: nothing carried in)
        if rank == 0:
            # rank 0 checks from the burst the cold checker is in step
            edge_ok = int(any(T.records_equal(want[(want["slot_bit"] >= 510 * j) & (want["slot_bit"] < hi)],
                                              got[(got["slot_bit"] >= 510 * j) & (got["slot_bit"] < hi)])[0] for j in range(1, 40)))
    ev = torch.tensor([edge_ok], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(ev, op=dist.ReduceOp.MIN)
    edge_ok = bool(int(ev[0]))
    # broadcast rank 0's single-GPU digest and compare
    ref = torch.tensor([one_digest & 0xffffffff, one_digest >> 32, ns_one], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.broadcast(ref, 0)
    ref = [int(x) for x in ref.cpu()]
    one_digest, ns_one = ref[0] | (ref[1] << 32), ref[2]
    ok = True
    for name in rec:
        same = rec[name].pop("_digest") == one_digest and rec[name].pop("_n") == ns_one
        rec[name]["digest_equals_single_gpu_decode"] = same
        ok = ok and same
    dd.close()
    del s_slots, s_pk
    first = next(iter(rec.values()))
    return {"workload": f"config5: rank 0's config-4 stream of {N_BURSTS} bursts (one bit per byte, 51 GB) sharded over {world} GPU(s) by "
                        "tb200_dist_rx_stream: packed on the device to 8 bits per byte, shards moved by grouped ncclSend/ncclRecv (or read in "
                        "place over NVLink: peer_packed), one all-gather of 56-byte summaries for the cell state, results rank-local",
            "scaling": "strong", "n_gpus": world, "value": first["value"], "unit": UNIT, "ms_per_step": first["ms_per_step"],
            "parity_ok": bool(ok and edge_ok), "single_gpu_slots": ns_one,
            "parity": {"digest_equals_single_gpu_decode": bool(ok), "shard_edge_windows_vs_cpu_checker": {"windows": world, "ok": edge_ok}},
            "modes": rec,
            "bound": "N=1: the decode pass (integer issue); N>1: rank 0 - it reads the whole 51 GB stream once to pack it (HBM) while it "
                     "decodes its own shard, the other ranks wait for their first chunk"}


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import tetra_testlib as T
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    G.load_package().load_library()          # fails loudly if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peaks, kc = Peaks(), kernel_constants()
    g = T.B200(device=local_rank)
    int_peak = g.lib.tb200_measure_int_peak(g.h)
    n = args.bursts
    shape = "config4"
    lead = SHAPES[shape]["lead_in_bits"]
    nbits = 510 * n + lead
    cfg = gen_cfg(T, rank_seed(rank), shape)
    buf = T.DevBuffer(g, nbits + 64)           # exportable memory: config 5's peer mode maps rank 0's stream into the other ranks
    d_bits = buf.tensor(torch.device("cuda", local_rank))
    rc = g.lib.tb200_gen_stream_dev(g.h, C.byref(cfg), 0, n, C.c_void_p(d_bits.data_ptr()), 1)
    assert rc == 0, g.err()
    ms = n + 16
    d_slots = torch.zeros(ms * 16, dtype=torch.uint8, device="cuda")
    d_t1 = torch.zeros(ms * 288, dtype=torch.uint8, device="cuda")
    g.set_options(chunk_bits=64, viterbi=T.VITERBI_LANE, output=T.OUT_UNPACKED, pipeline_slots=0, profile=1, input=T.IN_BYTES)

    def step_dev():
        ns = g.lib.tb200_rx_stream_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits, 3, C.c_void_p(d_slots.data_ptr()),
                                       C.c_void_p(d_t1.data_ptr()), None, ms)
        assert ns > 0.99 * n, (ns, g.err())
        return ns

    # ---- device-resident: value + per-kernel device times (CUDA events on the launching streams)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    tim = dict.fromkeys(TIM_KEYS, 0.0)
    launches = 0
    for _ in range(args.steps):
        ns = step_dev()
        t = g.timing()
        add_timing(tim, t)
        launches += g.stats().kernel_launches
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0            # this rank's own clock: t0 .. its last kernel done; the max over ranks is the job's
    barrier()
    st = g.stats()
    # the same steps once more with nothing overlapping (pass 1 on the decode stream): the kernels' own durations for the rooflines
    # (in the timed region pass 1 of piece i+1 runs under the decode pass of piece i and the event intervals include the waiting)
    tim_serial = serial_kernel_times(g, step_dev, max(2, min(args.steps, 5)))
    kinds = kinds_of(torch, d_slots, ns)
    dev_digest = None

    # ---- parity of the TIMED output: windows through the reference's own C + digest
    rx = None
    T.ensure_oracle_built()
    if not args.no_parity:
        rx = T.Ref() if T.have_ref() else T.Oracle()
    rng = np.random.default_rng(20261017 + rank)
    parity = None
    if rx is not None:
        if st.lock_losses == 0:
            parity = window_parity(T, torch, g, rx, d_bits, nbits, lead, d_slots, d_t1, ns, shape, args.parity_windows, rng)
        else:
            parity = {"checked": 0, "mismatches": 0, "note": f"{st.lock_losses} lock loss(es) in the run (a SYNC pattern at a wrong offset): window replay skipped"}
        parity["checker"] = "oracle/_ref (reference lower MAC compiled in place, restated osmo_conv_decode)" if T.have_ref() else "oracle port"
        parity["tie_rule"] = "ties keep predecessor s>>1 (include/tetra_tie_rule.h); libosmocore's own behaviour on noisy input is unpinned"
        if parity["mismatches"]:
            raise SystemExit(f"bench: the timed output differs from the reference: {parity}")

    # ---- end to end: pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        import psutil
        barrier()                                   # every rank looks at the host's memory before any of them takes its share
        avail = psutil.virtual_memory().available / max(1, world)
        barrier()
        n_e2e = n
        need = lambda nb: (510 * nb + lead) + (nb + 16) * (16 + 288)
        while need(n_e2e) > 0.5 * avail and n_e2e > 1_000_000:
            n_e2e //= 2
        # pinned memory is a scarcer thing than "available" says (eight ranks allocate at once): halve until it fits
        while True:
            nbits_e = 510 * n_e2e + lead
            ms_e = n_e2e + 16
            h_bits_p = g.lib.tb200_host_alloc(nbits_e)
            h_slots_p = g.lib.tb200_host_alloc(ms_e * 16)
            h_t1_p = g.lib.tb200_host_alloc(ms_e * 288)
            if h_bits_p and h_slots_p and h_t1_p:
                break
            for p_ in (h_bits_p, h_slots_p, h_t1_p):
                if p_:
                    g.lib.tb200_host_free(p_)
            assert n_e2e > 250_000, "pinned host allocation failed"
            n_e2e //= 2
        print("bench: rank %d e2e leg with %d bursts (host memory available per rank %.1f GB)" % (rank, n_e2e, avail / 1e9), file=sys.stderr)
        h_bits = torch.from_numpy(np.ctypeslib.as_array(C.cast(h_bits_p, C.POINTER(C.c_uint8)), shape=(nbits_e,)))
        h_bits.copy_(d_bits[:nbits_e])
        g.set_options(profile=0, output=T.OUT_UNPACKED)

        def step_host():
            k = g.lib.tb200_rx_stream_host(g.h, h_bits_p, nbits_e, 3, h_slots_p, h_t1_p, None, ms_e)
            assert k > 0.99 * n_e2e, (k, g.err())
            return k
        e_steps = args.steps if n_e2e < 20_000_000 else min(args.steps, args.e2e_steps)
        for _ in range(2):
            step_host()
        barrier()
        t1 = time.perf_counter()
        for _ in range(e_steps):
            ns_e = step_host()
        wall_e2e = time.perf_counter() - t1
        barrier()
        # the host path delivered what the device path did (whole run when it ran at full size)
        hs = torch.from_numpy(np.ctypeslib.as_array(C.cast(h_slots_p, C.POINTER(C.c_uint8)), shape=(ns_e * 16,)))
        ht = torch.from_numpy(np.ctypeslib.as_array(C.cast(h_t1_p, C.POINTER(C.c_uint8)), shape=(ns_e * 288,)))
        same = True
        if n_e2e == n:
            CH = 4_000_000
            for a in range(0, ns_e, CH):
                b = min(ns_e, a + CH)
                same = same and torch.equal(hs[a * 16:b * 16].cuda(), d_slots[a * 16:b * 16]) and torch.equal(ht[a * 288:b * 288].cuda(), d_t1[a * 288:b * 288])
        e2e = {"wall": wall_e2e, "steps": e_steps, "slots": ns_e, "bursts": n_e2e, "same": bool(same), "h2d": nbits_e, "d2h": ns_e * (16 + 288)}
        for p in (h_bits_p, h_t1_p):
            g.lib.tb200_host_free(p)
        # the same call with both sides bit-packed (TB200_IN_PACKED in, TB200_OUT_PACKED out): 8x fewer bytes over PCIe
        nb8 = 4 * ((nbits_e + 31) // 32)
        d_pk_in = torch.zeros(nb8 + 64, dtype=torch.uint8, device="cuda")
        assert g.lib.tb200_pack_bits_dev(g.h, C.c_void_p(d_bits.data_ptr()), nbits_e, C.c_void_p(d_pk_in.data_ptr())) == 0, g.err()
        h_pk_in_p = h_pk_out_p = None
        for attempt in range(8):                   # the pages the other ranks just released may not be back yet
            h_pk_in_p = g.lib.tb200_host_alloc(nb8 + 64)
            h_pk_out_p = g.lib.tb200_host_alloc(ms_e * 36)
            if h_pk_in_p and h_pk_out_p:
                break
            for p_ in (h_pk_in_p, h_pk_out_p):
                if p_:
                    g.lib.tb200_host_free(p_)
            h_pk_in_p = h_pk_out_p = None
            time.sleep(0.5)
        ok_all = torch.tensor([1 if (h_pk_in_p and h_pk_out_p) else 0], dtype=torch.int32, device="cuda")
        if dist is not None:
            dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        if int(ok_all[0]) == 0:                     # one rank could not pin its buffers: every rank leaves the packed leg out
            print("bench: rank %d: no pinned memory for the bit-packed host leg (%.1f GB available), leg skipped" % (
                rank, psutil.virtual_memory().available / 1e9), file=sys.stderr)
            e2e["wall_packed"] = 0.0
            e2e["h2d_packed"] = e2e["d2h_packed"] = 0
            e2e["same_packed"] = None
            del d_pk_in
        else:
            torch.from_numpy(np.ctypeslib.as_array(C.cast(h_pk_in_p, C.POINTER(C.c_uint8)), shape=(nb8,))).copy_(d_pk_in[:nb8])
            del d_pk_in
            g.set_options(input=T.IN_PACKED, output=T.OUT_PACKED)

            def step_host_packed():
                k = g.lib.tb200_rx_stream_host(g.h, h_pk_in_p, nbits_e, 3, h_slots_p, None, h_pk_out_p, ms_e)
                assert k == ns_e, (k, g.err())
            step_host_packed()
            barrier()
            t2 = time.perf_counter()
            for _ in range(e_steps):
                step_host_packed()
            e2e["wall_packed"] = time.perf_counter() - t2
            e2e["h2d_packed"], e2e["d2h_packed"] = nb8, ns_e * (16 + 36)
            barrier()
            hs2 = torch.from_numpy(np.ctypeslib.as_array(C.cast(h_slots_p, C.POINTER(C.c_uint8)), shape=(ns_e * 16,)))
            e2e["same_packed"] = bool(torch.equal(hs2[:16 * 4_000_000].cuda(), d_slots[:16 * 4_000_000])) if n_e2e == n else True
            g.set_options(input=T.IN_BYTES, output=T.OUT_UNPACKED)
        for p in (h_slots_p, h_pk_in_p, h_pk_out_p):
            if p:
                g.lib.tb200_host_free(p)
        if not same:
            raise SystemExit("bench: the host-buffer path and the device-resident path delivered different output")
    clocks = sampler.stop() if rank == 0 else None

    # ---- the stand-alone descramble + de-interleave stage (north star: >= 70 % of the HBM roofline)
    g.set_options(profile=1)
    del d_t1
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
    del d5, d3, dcodes, d_slots

    # ---- config 5: rank 0's stream over all ranks
    config5 = None
    if not args.no_config5:
        config5 = run_config5(g, T, torch, dist, rank, world, d_bits, nbits, rx, rng, args.config5_steps, 2)

    # ---- the other configurations as named sub-records (N = 1 only)
    configs = None
    if world == 1 and not args.no_configs:
        del d_bits
        buf.free()
        configs = {"config2": run_shape(g, T, torch, "config2", 1_000_000, 20, 5, 0x7E7A0002, int_peak, peaks, kc, rx, rng),
                   "config3": run_shape(g, T, torch, "config3", 10_000_000, 10, 3, 0x7E7A0003, int_peak, peaks, kc, rx, rng)}

    vals = [wall, tim_serial["total"], tim_serial["classify"], tim_serial["scan"], tim_serial["decode"], tim_serial["search"]] + \
           ([e2e["wall"]] if e2e else [0.0]) + [tim["total"]] + ([e2e["wall_packed"]] if e2e else [0.0]) + [tim_serial["prepare"], tim_serial["trellis"]]
    wall, t_total, t_cls, t_scan, t_dec, t_search, wall_e2e, t_total_overlapped, wall_e2e_packed, t_prep, t_trel = reduce_max(dist, vals, "cuda")
    tot_slots = torch.tensor([ns, e2e["slots"] if e2e else 0], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot_slots)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    tim_max = {"total": t_total, "classify": t_cls, "scan": t_scan, "decode": t_dec, "search": t_search, "prepare": t_prep, "trellis": t_trel}
    dec, search, share = rooflines(kinds, ns, tim_max, 1, int_peak, peaks, kc, shape)
    dec["timing"] = ("kernel durations: CUDA events on the launching streams over steps run with options.serial_passes = 1 (no two kernels "
                     "side by side); the timed region of `value` overlaps pass 1 of piece i+1 with the decode pass of piece i")
    stage_gbs = 2 * 432 * nblk / (stage_ms * 1e-3) / 1e9
    dec["peak_source"] = "tb200_measure_int_peak (this run, this GPU): %.3g integer thread-instructions/s" % int_peak
    dec["sync_search"] = search
    dec["descramble_deinterleave_stage"] = {"kernel": "k_stage_tma", "bound": "hbm", "achieved": stage_gbs, "peak": peaks.hbm, "unit": "GB/s",
                                            "frac": stage_gbs / peaks.hbm, "ms_per_launch": stage_ms, "algorithmic_bytes_per_block": 864, "blocks": nblk}
    dec["step_share"] = share
    dec["constants_from"] = kc.get("source")
    value = int(tot_slots[0]) * args.steps / wall
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bursts_per_gpu_per_step": n, "slots_decoded_per_gpu": int(ns),
                       "kinds": {"dropped": kinds[0], "sync": kinds[1], "ndb_schf": kinds[2], "ndb_two_blocks": kinds[3]},
                       "lock_losses": int(st.lock_losses), "crc_ok_blocks": int(st.crc_ok_blocks), "blocks": int(st.blocks),
                       "viterbi": "lane: two packed trellises per thread", "l2": "inputs larger than L2 (51 GB stream per step)",
                       "parallelism": f"independent streams x{world}", "output": "slot records + unpacked type-1 bits (1 bit/byte)",
                       "host_numa": numa},
            "gpu_launches": int(launches), "device_ms_per_step": t_total_overlapped / args.steps, "device_ms_per_step_serial_passes": t_total,
            "roofline": dec, "clocks": clocks}
    if e2e:
        line["e2e"] = {"value": int(tot_slots[1]) * e2e["steps"] / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"] * world,
                       "d2h_bytes_per_step": e2e["d2h"] * world, "ms_per_step": wall_e2e / e2e["steps"] * 1e3, "steps": e2e["steps"],
                       "bursts_per_gpu_per_step": e2e["bursts"], "matches_device_path": e2e["same"],
                       "note": "tb200_rx_stream_host, pinned host buffers, one byte per bit in and out (the reference's ABI on both sides): "
                               "bound by PCIe, and beyond four GPUs by the host's aggregate path to its GPUs (one NUMA node visible in this VM, nothing to bind to)",
                       }
        if wall_e2e_packed > 0:
            line["e2e"]["packed_io"] = {"value": int(tot_slots[1]) * e2e["steps"] / wall_e2e_packed, "unit": UNIT,
                                        "h2d_bytes_per_step": e2e["h2d_packed"] * world, "d2h_bytes_per_step": e2e["d2h_packed"] * world,
                                        "ms_per_step": wall_e2e_packed / e2e["steps"] * 1e3, "matches_device_path": e2e["same_packed"],
                                        "note": "the same call with TB200_IN_PACKED input and TB200_OUT_PACKED output: eight bits per byte both ways"}
    if parity:
        line["parity"] = parity
    if config5:
        line["config5"] = config5
    if configs:
        line["configs"] = configs
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_reference_rate(20000, os.cpu_count() or 1, one_core=True)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bursts", type=lambda v: int(float(v)), default=N_BURSTS, help="bursts of the headline stream (profiling runs use less)")
    ap.add_argument("--e2e-steps", type=int, default=5, help="steps of the host-buffer leg when it runs at full size (1 s per step)")
    ap.add_argument("--config5-steps", type=int, default=5)
    ap.add_argument("--parity-windows", type=int, default=12)
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--no-cpu", action="store_true", help="profiling runs: skip the CPU baseline leg")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
