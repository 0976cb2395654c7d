"""profiles/traffic.json from an `ncu -i X.ncu-rep --page raw --csv` dump: DRAM bytes (read + write) per launch of each
kernel -- what bench.py reports as roofline.traffic.   usage: ncu_traffic.py raw.csv "<capture name>" """
import csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in body:
    name = r[ik].replace("void ", "").split("(NmfScene")[0].split("(const")[0].replace("(int)", "").replace(", ", ",")
    b = float(r[ir].replace(",", "")) * mul[units[ir]] + float(r[iw].replace(",", "")) * mul[units[iw]]
    out.setdefault(name, b)
    if name.startswith(("k_march<", "k_shade<")) and name.endswith(",0>"):      # <LEVEL,TRAIN=0>: bench.py asks by level
        out.setdefault(name[:-3] + ">", b)
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump({"capture": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": out}, open(p, "w"), indent=1)
print(json.dumps(out, indent=1))
