"""Join an ncu launch list (csv) with the B2P_DEBUG gemm shape log: per-GEMM TFLOP/s and share of the forward."""
import csv, sys, collections
csv_path, log_path = sys.argv[1], sys.argv[2]
n_fwd = int(sys.argv[3]) if len(sys.argv) > 3 else 241
rows = list(csv.reader(open(csv_path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
def us(r):
    v = float(r[vi].replace(',', ''))
    return v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
names = [r[ki] for r in data]
start = next(i for i, n in enumerate(names) if 'im2col' in n)   # first launch of a forward
after = [us(r) for r in data[start:] if 'gemm' in r[ki]]
before = [us(r) for r in data[:start] if 'gemm' in r[ki]]
other = sum(us(r) for r in data if 'gemm' not in r[ki] and any(k in r[ki] for k in ('adown', 'cbfuse', 'im2col', 'upsample', 'maxpool')))
shapes = []
for line in open(log_path):
    if line.startswith('b2p_gemm'):
        d = dict(kv.split('=') for kv in line.split()[1:])
        shapes.append({k: int(v) for k, v in d.items()})
shapes = shapes[:n_fwd]
pairs = list(zip(after, shapes[:len(after)])) + list(zip(before, shapes[n_fwd - len(before):]))
gem = [p[0] for p in pairs]
shapes = [p[1] for p in pairs]
print(f"{len(gem)} of {n_fwd} gemm launches of one forward captured")
tot = sum(gem) + other
print(f"total {tot:.0f} us, gemm {sum(gem):.0f} us, other {other:.0f} us")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for t, s in zip(gem, shapes):
    fl = 2.0 * s['M'] * s['N'] * s['Ktot']
    key = (s['mode'], s['M'], s['N'], s['Ktot'], s['bn'], s['tw'], s['th'], s['m_tiles'] * s['n_tiles'], s['stages'], s['act'], s['res'])
    a = agg[key]; a[0] += 1; a[1] += t; a[2] += fl
print("mode      M     N  Ktot  bn  tw th tiles st act res |  n   us_tot  us_each  TFLOP/s  %time")
for key, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%4d %7d %5d %5d %3d %3d %2d %5d %2d %3d %3d | %2d %8.1f %8.1f %8.1f %6.1f" % (*key, n, t, t / n, fl / t / 1e6, 100 * t / tot))
