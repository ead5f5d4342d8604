"""per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv).  usage: python tools/launch_summary.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= iV: continue
    v = float(r[iV].replace(',', '')); u = r[iU]
    ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else (v if u.startswith('m') else v * 1e3))
    k = r[iK].split('(')[0]
    agg[k][0] += 1; agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} launches   total ms   share")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {c:6d} {ms:10.2f} {ms / tot * 100:6.1f} %")
