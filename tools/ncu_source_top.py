"""Top source lines by warp-stall samples from `ncu -i rep --page source --csv --print-source cuda,sass --kernel-id :::N`.
  ncu -i gpurun_out/x.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::4 > /tmp/s.csv; python tools/ncu_source_top.py /tmp/s.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Line No")
si = hdr.index("# Samples")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
ie = hdr.index("Instructions Executed")
cur_file = ""
data = []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if len(r) > si and r[0].isdigit() and r[si].isdigit():
        data.append((int(r[si]), cur_file, r[0], r[1].strip()[:100], sorted([(int(r[i]), hdr[i][6:]) for i in stall if r[i].isdigit() and int(r[i]) > 0], reverse=True)[:3], r[ie]))
tot = sum(d[0] for d in data)
print("total samples", tot)
for d in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{100 * d[0] / tot:5.1f}% {d[1]}:{d[2]:>4s} inst={d[5]:>9s} {d[3]:100s} {d[4]}")
