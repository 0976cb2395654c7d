"""Top SASS instructions by stall samples from `ncu --page source --csv` output.  usage: ncu_sass_top.py file.csv [N] [kernel substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    name = rows[a][1]
    if want not in name:
        continue
    hdr = rows[a + 1]
    body = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    cs, ci, cx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[cs] or 0) for r in body)
    print("==", name, "total samples", tot, "instructions", len(body))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][cs] or 0))[:N]
    for i in sorted(order):
        r = body[i]
        top = sorted(((int(r[j] or 0), hdr[j][6:]) for j in stalls), reverse=True)[:2]
        print(f"{i:5d} {int(r[cs]):7d} {100*int(r[cs])/tot:5.1f}% x{r[cx]:>10s}  {r[ci][:64]:64s} {top}")
