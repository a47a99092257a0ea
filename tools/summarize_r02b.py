"""gpurun_out/{prof_r02b_config*.ncu-rep, r02b_launches.csv, run log with the LIVE lines} -> profiles/r02b_summary.md,
profiles/r02b_launches.csv, profiles/kernel_constants.json (per-slot / per-ACS constants that bench.py quotes),
profiles/r02b_acs_loop.sass.   usage: python tools/summarize_r02b.py <run-log>"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out, src = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")
runlog = open(sys.argv[1]).read() if len(sys.argv) > 1 else ""
ACS_SB1, ACS_HALF, ACS_SCHF = 16 * 84, 16 * 148, 16 * 292
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
lines = ["# ncu summary r02b (final tree of round 2: split decode pass)", "",
         "`ncu --set full --clock-control none --import-source on -k regex:'k_lane_prepare|k_lane_trellis|k_classify_tile|k_sb1_lane' -s 4 -c 4 python tools/prof_run.py <shape> 2097152 2`",
         "(one launch each over a piece of 2^21 slots, options.serial_passes = 1; cold-cache serialised replays: compare shares and counts, not absolute times;",
         "the LIVE lines are the same script's CUDA-event times without the profiler)", ""]
consts = {"source": "profiles/r02b_summary.md (ncu --set full, one launch over 2^21 slots of the named shape, tools/prof_run.py)"}
for shape in ("config4", "config2"):
    rep = os.path.join(src, f"prof_r02b_{shape}.ncu-rep")
    if not os.path.exists(rep):
        continue
    log = open(os.path.join(src, f"ncu_r02b_{shape}.log")).read()
    m = re.search(r"PROF \S+ slots (\d+) kinds \[(\d+), (\d+), (\d+), (\d+)\]", log)
    ns, kinds = int(m.group(1)), [int(m.group(i)) for i in range(2, 6)]
    live = [l for l in runlog.splitlines() if l.startswith(f"LIVE {shape}")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines += [f"## {shape}: {ns} slots, kinds dropped/SB/SCH-F/two-block = {kinds}", ""]
    if live:
        lines += ["CUDA-event times of the same launches without the profiler: `" + live[-1] + "`", ""]
    acs_dec = kinds[1] * ACS_HALF + kinds[2] * ACS_SCHF + kinds[3] * 2 * ACS_HALF
    acs = {"k_lane_trellis": acs_dec, "k_decode_lane": acs_dec, "k_sb1_lane": kinds[1] * ACS_SB1}
    for r in rows[2:]:
        kname = re.sub(r"^void ", "", r[idx["Kernel Name"]]).split("<")[0].split("(")[0].replace("tb::", "")
        lines += [f"### {kname} ({shape})", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                lines.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        st = []
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and r[idx[k]] != "":
                st.append((float(r[idx[k]]), k))
        for v, k in sorted(st, reverse=True)[:6]:
            lines.append(f"| {k} | {v:.2f} | warps/issue |")
        dram = float(r[idx["dram__bytes_read.sum"]]) * UNIT[units[idx["dram__bytes_read.sum"]]] + \
            float(r[idx["dram__bytes_write.sum"]]) * UNIT[units[idx["dram__bytes_write.sum"]]]
        tinst = float(r[idx["smsp__inst_executed.sum"]]) * 32
        c = {"dram_bytes_per_slot": dram / ns, "thread_inst_per_slot": tinst / ns, "ncu_duration_us": float(r[idx["gpu__time_duration.sum"]]), "slots": ns}
        if acs.get(kname):
            c["thread_inst_per_acs"] = tinst / acs[kname]
            lines.append(f"| thread instructions per add-compare-select ({acs[kname]} ACS in the launch) | {c['thread_inst_per_acs']:.3f} | |")
        lines.append(f"| DRAM traffic per slot (read + write) | {dram / ns:.1f} | byte |")
        lines.append("")
        consts.setdefault(kname, {})[shape] = c
for k in list(consts):
    if isinstance(consts[k], dict) and "config2" in consts[k] and "config3" not in consts[k]:
        consts[k]["config3"] = dict(consts[k]["config2"], note="config-2 capture (same kernels, mostly SCH/F)")
json.dump(consts, open(os.path.join(out, "kernel_constants.json"), "w"), indent=1)
lc = os.path.join(src, "r02b_launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 14 and r[0].isdigit()]
    tot, cnt = {}, {}
    for r in rows:
        name = re.sub(r"^void ", "", r[4]).split("<")[0].split("(")[0]
        tot[name] = tot.get(name, 0.0) + float(r[-1]); cnt[name] = cnt.get(name, 0) + 1
    alls = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith("tb::k_") and "gen" not in k)
    lines += ["## launch list of a short bench run", "",
              "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 --bursts 8000000 --no-e2e --no-cpu --no-config5 --no-configs --no-parity`",
              "(first 400 launches; raw list: `profiles/r02b_launches.csv`; `share of chain` = among the receive chain's own kernels, i.e. without the stream generator and the peak measurement)",
              "", "| kernel | launches | total ns | share | share of chain |", "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        chain = f"{v / ours:.3f}" if (k.startswith("tb::k_") and "gen" not in k) else ""
        lines.append(f"| {k} | {cnt[k]} | {v:.0f} | {v / alls:.3f} | {chain} |")
    open(os.path.join(out, "r02b_launches.csv"), "w").write(open(lc).read())
open(os.path.join(out, "r02b_summary.md"), "w").write("\n".join(lines) + "\n")
r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "acs_loop_mix.py"), os.path.join(ROOT, "osmo-tetra_b200", "libtetra_b200.so"),
                    "k_lane_trellisILb0ELb1", "/tmp/acs_loop.sass"], capture_output=True, text=True).stdout.strip()
head = ["// k_lane_trellis<false, true>: the ACS loop of viterbi_pair_u8 (eight trellis steps of two packed trellises per iteration)",
        "// " + r, "// cuobjdump -sass osmo-tetra_b200/libtetra_b200.so (sm_100a), extracted by tools/acs_loop_mix.py", ""]
open(os.path.join(out, "r02b_acs_loop.sass"), "w").write("\n".join(head) + open("/tmp/acs_loop.sass").read())
print(r)
print(json.dumps(consts, indent=1)[:2500])
