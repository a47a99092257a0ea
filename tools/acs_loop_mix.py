"""instruction mix of the ACS loop (eight-step form) inside a kernel of a built library: python tools/acs_loop_mix.py <lib.so> [mangled-name-substring]"""
import re, subprocess, sys
lib = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else "k_lane_trellisILb0ELb1"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
m = re.search(r"Function : (\S*" + re.escape(key) + r"\S*)(.*?)\n\s*\.\.\.\.\.\.\.\.\.\.", sass, re.S)
body = m.group(2)
ins = [(int(x.group(1), 16), x.group(0)) for x in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+[^\n]*", body)]
best = None
for addr, text in ins:
    b = re.search(r"BRA\s+(?:U\w+,\s*)?0x([0-9a-f]+)", text)
    if b and int(b.group(1), 16) < addr:
        lo = int(b.group(1), 16)
        seg = [t for a, t in ins if lo <= a <= addr]
        if sum("VIADDMNMX" in t for t in seg) >= 120 and (best is None or len(seg) < len(best)):
            best = seg
ALU = {"VIADDMNMX", "LOP3", "VIADD", "PRMT", "SHF", "IADD3", "VIMNMX", "ISETP", "SEL", "LEA", "MOV", "BMSK", "SGXT", "POPC", "FLO", "IABS"}
ops = {}
for t in best:
    op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", t).group(1)
    k = ".".join(op.split(".")[:2]) if op.startswith("IMAD") and op.split(".")[1:2] and op.split(".")[1] in ("IADD", "MOV", "SHL", "U32", "WIDE") else op.split(".")[0]
    ops[k] = ops.get(k, 0) + 1
alu = sum(v for k, v in ops.items() if k in ALU); fma = sum(v for k, v in ops.items() if k.startswith("IMAD"))
print(m.group(1)[:60], len(best), "instructions; ALU", alu, "FMA", fma, "|", ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(best) + "\n")
