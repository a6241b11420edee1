"""Per-kernel and per-GEMM-shape table of one profiled parse step: python tools/step_table.py <launches.csv> <shapes.log>"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
def us(r):
    v = float(r[vi].replace(',', '')); return v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in data:
    n = r[ki].split('(')[0].replace('b2p::', '').replace('void ', '')
    tot[n] += us(r); cnt[n] += 1
T = sum(tot.values()); print("launches", len(data), "total us", round(T))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:16]:
    print(f"{v:10.1f} us {100 * v / T:5.1f}% n={cnt[k]:5d} avg {v / cnt[k]:8.1f}  {k[:80]}")
shapes = []; started = False
for line in open(sys.argv[2]):
    if 'PROFILE_START' in line: started = True; continue
    if 'PROFILE_STOP' in line: break
    if started and line.startswith('b2p_gemm'):
        d = dict(kv.split('=') for kv in line.split()[1:]); shapes.append({k: int(v) for k, v in d.items()})
gem = [us(r) for r in data if 'gemm_tcgen05' in r[ki]]
assert len(gem) == len(shapes), (len(gem), len(shapes))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for t, s in zip(gem, shapes):
    key = (s['mode'], s['M'], s['N'], s['Ktot'], s['bn'], s.get('ksplit', 1), s['m_tiles'] * s['n_tiles'], s['stages'], s['act'], s['f32'], s['res'], s.get('x3', 0))
    a = agg[key]; a[0] += 1; a[1] += t; a[2] += 2.0 * s['M'] * s['N'] * s['Ktot'] * (3 if s.get('x3', 0) else 1)
print("mode      M     N   Ktot  bn ks tiles st act f32 res x3 |   n   us_tot  us_each TFLOP/s(exec) %gemm")
G = sum(gem)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
for key, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%4d %7d %5d %6d %3d %2d %5d %2d %3d %3d %3d %2d | %3d %8.1f %8.1f %8.1f %6.1f" % (*key, n, t, t / n, fl / t / 1e6, 100 * t / G))
