"""Turn gpurun_out/{prof.ncu-rep,launches.csv,bench.json} into the tracked summaries under profiles/.
usage: python tools/summarize_profile.py <tag>      (e.g. r01b)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"# ncu summary {tag}", ""]
rep = os.path.join(src, "prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines.append("`ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_classify' python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu`")
    lines.append("(10^6-burst config-2 stream per launch; cold-cache serialised replays: compare shares, not absolutes)")
    consts = {"source": f"profiles/{tag}_summary.md (ncu --set full, 10^6-burst config-2 launch)"}
    def _bytes(v, unit):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    for r in rows[2:]:
        kname = r[idx['Kernel Name']].split("(")[0].split("::")[-1].replace("void ", "").split("<")[0].strip()
        if float(r[idx["gpu__time_duration.sum"]]) > consts.get(kname, {}).get("_dur", 0):
            consts[kname] = {"_dur": float(r[idx["gpu__time_duration.sum"]]),
                             "dram_bytes_per_launch": _bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
                                                      _bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]),
                             "thread_inst_per_launch": float(r[idx["smsp__inst_executed.sum"]]) * 32,
                             "ncu_duration_us": float(r[idx["gpu__time_duration.sum"]])}
    for v in consts.values():
        if isinstance(v, dict):
            v.pop("_dur", None)
    json.dump(consts, open(os.path.join(out, "kernel_constants.json"), "w"), indent=1)
    for r in rows[2:]:
        lines += ["", f"## {r[idx['Kernel Name']]}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                lines.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(r[idx[k]]) >= 0.15:
                        lines.append(f"| {k} | {r[idx[k]]} | warps/issue |")
                except ValueError:
                    pass
        rd, wr = r[idx["dram__bytes_read.sum"]], r[idx["dram__bytes_write.sum"]]
        lines.append(f"| traffic (dram read + write per launch) | {rd} + {wr} | {units[idx['dram__bytes_read.sum']]} |")
lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    lines += ["", "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, first step shown)", "",
              "| # | kernel | grid | block | ns |", "|---|---|---|---|---|"]
    rows = [r for r in csv.reader(open(lc)) if len(r) > 14 and r[0].isdigit()]
    tot = {}
    for r in rows:
        name = r[4].split("(")[0]
        tot[name] = tot.get(name, 0) + float(r[14])
    for r in rows[:26]:
        lines.append(f"| {r[0]} | {r[4].split('(')[0][:60]} | {r[8]} | {r[7]} | {r[14]} |")
    ours = {k: v for k, v in tot.items() if "tb::" in k or k.startswith("k_") or "k_classify" in k}
    s = sum(ours.values())
    lines += ["", "share of our kernels over the captured launches:", ""]
    for k, v in sorted(ours.items(), key=lambda x: -x[1]):
        lines.append(f"- {k}: {v/1e6:.3f} ms ({100*v/s:.1f} %)")
    with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
        f.write(open(lc).read())
for name in ("bench.json", "bench_ref.json"):
    p = os.path.join(src, name)
    if os.path.exists(p):
        for l in open(p):
            if l.strip().startswith("{"):
                lines += ["", f"## {name}", "", "```json", json.dumps(json.loads(l), indent=1), "```"]
open(os.path.join(out, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("wrote", os.path.join(out, f"{tag}_summary.md"))
