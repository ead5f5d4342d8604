"""rank CUDA source lines of an ncu report (--import-source on, -lineinfo) by shared-memory wavefronts (and the excess due to
bank conflicts).  usage: python tools/ncu_smem_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]; iw = h.index("L1 Wavefronts Shared"); ie = h.index("L1 Wavefronts Shared Excessive"); ii = h.index("L1 Wavefronts Shared Ideal")
agg = {}; cur = ""
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]
    if r and r[0].isdigit() and len(r) > iw:
        try: agg[(cur, int(r[0]), r[1])] = (float(r[iw] or 0), float(r[ie] or 0), float(r[ii] or 0))
        except ValueError: pass
tot = sum(v[0] for v in agg.values()) or 1
print("total shared-memory wavefronts", int(tot), "excessive", int(sum(v[1] for v in agg.values())))
for (f, ln, src), (w, e, i_) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} wavefronts={w / tot * 100:5.1f}% excessive={e / max(w, 1) * 100:5.1f}%  {src.strip()[:100]}")
