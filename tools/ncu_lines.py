"""rank CUDA source lines of an ncu report (captured with --import-source on, -lineinfo) by stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
agg = {}; cur = ""
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]
    if r and r[0].isdigit() and len(r) > iI:
        try: agg[(cur, int(r[0]), r[1])] = (int(r[iI]), int(r[iS]))
        except ValueError: pass
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print("total warp instructions", ti, "samples", ts)
for (f, ln, src), (ins, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:4d} inst={ins / ti * 100:5.1f}% samp={sm / ts * 100:5.1f}%  {src.strip()[:100]}")
