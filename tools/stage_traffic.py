"""Per-stage DRAM traffic / time of ONE profiled parse step from an ncu launch list with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (tools/profile_step.py, eager launches):
  python tools/stage_traffic.py gpurun_out/step_traffic.csv > profiles/r2_stage_traffic.json
Stages are cut at kernel names: detect = letterbox .. batched_nms; crop = crop_resize; encode = up to the first decoder_embed;
decode_step = decoder_embed .. step_advance (averaged over the steps)."""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ii, ki, mi, vi, ui = cols.index("ID"), cols.index("Kernel Name"), cols.index("Metric Name"), cols.index("Metric Value"), cols.index("Metric Unit")
launch = collections.OrderedDict()
for r in data:
    d = launch.setdefault(int(r[ii]), {"name": r[ki].split("(")[0].replace("b2p::", "").replace("void ", "")})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi].startswith("gpu__time"):
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d["rd" if "read" in r[mi] else "wr"] = v * scale
stage, steps = "detect", 0
out = collections.defaultdict(lambda: {"us": 0.0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "launches": 0})
for d in launch.values():
    n = d["name"]
    if n.startswith("crop_resize"):
        stage = "crop"
    elif stage == "crop" and not n.startswith("crop_resize"):
        stage = "encode"
    if n.startswith("decoder_embed"):
        stage = "decode"
        steps += 1
    o = out[stage]
    o["us"] += d.get("us", 0.0); o["dram_bytes_read"] += d.get("rd", 0.0); o["dram_bytes_write"] += d.get("wr", 0.0); o["launches"] += 1
res = {}
for k, o in out.items():
    div = steps if k == "decode" and steps else 1
    name = "decode_step" if k == "decode" else k
    res[name] = {"ncu_us": o["us"] / div, "dram_bytes_read": o["dram_bytes_read"] / div, "dram_bytes_write": o["dram_bytes_write"] / div,
                 "dram_bytes": (o["dram_bytes_read"] + o["dram_bytes_write"]) / div, "launches": o["launches"] // div}
res["_note"] = f"one parse step, 8 screenshots, eager launches under ncu (caches flushed per kernel: upper bound of the live traffic); {steps} decode steps averaged"
print(json.dumps(res, indent=1))
