"""Per-kernel SASS opcode histograms of libnmf_b200.so (cuobjdump -sass), the evidence for "Blackwell-native" claims:
UTCHMMA / UTCBAR / LDTM (tcgen05 MMA, commit, TMEM loads), UBLKCP (TMA bulk copies), HMMA (legacy mma.sync), RED / ATOMG,
LDG.256 / STG.256 (256-bit global accesses, counted in addition to LDG / STG), UCGABAR_ARV (thread-block cluster barrier).
usage: python tools/sass_histogram.py [out_dir]   -> profiles/sass/<kernel>.txt + profiles/sass/SUMMARY.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass")
os.makedirs(out_dir, exist_ok=True)
so = os.path.join(ROOT, "nmf_b200", "libnmf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur][op.split(".")[0]] += 1
        if re.match(r"(LDG|STG)\..*\b256\b", op):                     # 256-bit global loads / stores (new on sm_100)
            kernels[cur][op.split(".")[0] + ".256"] += 1
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "HMMA", "RED", "REDG", "ATOMG", "ATOM", "LDG", "STG", "LDS", "STS",
       "SHFL", "FFMA", "MUFU", "BAR", "SYNCS", "LDG.256", "STG.256", "UCGABAR_ARV"]
rows = []
for mangled, c in sorted(kernels.items()):
    name = demangle(mangled)
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")
    with open(os.path.join(out_dir, fn + ".txt"), "w") as f:
        f.write(f"# {name}\n# SASS opcode histogram (cuobjdump -sass nmf_b200/libnmf_b200.so), {sum(c.values())} instructions\n")
        for op, n in c.most_common():
            f.write(f"{n:7d}  {op}\n")
    rows.append((short, sum(c.values()), [c.get(k, 0) for k in KEY]))
with open(os.path.join(out_dir, "SUMMARY.md"), "w") as f:
    f.write("# SASS opcode counts per kernel (sm_100a), from `python tools/sass_histogram.py`\n\n")
    f.write("| kernel | instr | " + " | ".join(KEY) + " |\n|---|---|" + "---|" * len(KEY) + "\n")
    for short, tot, vals in rows:
        f.write(f"| `{short}` | {tot} | " + " | ".join(str(v) if v else "" for v in vals) + " |\n")
print(open(os.path.join(out_dir, "SUMMARY.md")).read())
