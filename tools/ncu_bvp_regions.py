"""aggregate an ncu report of k_vl_bvp (captured with --import-source on) by code region of vl_bvp.cuh"""
import csv, subprocess, sys, os
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instructions", ti, "samples", ts)
src = open(os.path.join(root, "aeroapps_b200/csrc/vl_bvp.cuh")).read().splitlines()
def find(s): return [i + 1 for i, l in enumerate(src) if s in l][0]
marks = [("panel fn", find("DEVFN void bvp_panel(")), ("dispatch/urows fn", find("DEVFN void bvp_panel_dispatch")), ("bvp_pixel prologue", find("DEVFN void bvp_pixel(")),
         ("load new rows", find("// ---- load the new rows")), ("elim ctl", find("// ---- eliminate the 2n columns")), ("lookahead upd", find("// everyone: apply panel P")),
         ("panel call+trailing", find("if (is_panel && has_next) bvp_panel_dispatch")), ("barrier after trailing", find("        T.sync();\n") if False else find("if (iflag[0]) { failed = true; break; }\n        if (has_next)") if False else find("// trailing update: W(r, c)") + 36),
         ("spill/shift", find("double *Uk = Ug + (size_t)k * n2 * Cc;")), ("backward", find("// ---- back substitution fused")), ("tri solve", find("// upper-triangular solve")), ("source terms", find("// xcur = [LCON"))]
marks.sort(key=lambda x: x[1])
for i, (name, a) in enumerate(marks):
    b = marks[i + 1][1] - 1 if i + 1 < len(marks) else 10 ** 6
    ii = sum(v[0] for (f, l, _), v in agg.items() if f == "vl_bvp.cuh" and a <= l <= b); ss = sum(v[1] for (f, l, _), v in agg.items() if f == "vl_bvp.cuh" and a <= l <= b)
    print(f"{name:24s} lines {a}-{b}: inst={ii / ti * 100:5.1f}% samp={ss / ts * 100:5.1f}%")
ii = sum(v[0] for (f, l, _), v in agg.items() if f != "vl_bvp.cuh"); ss = sum(v[1] for (f, l, _), v in agg.items() if f != "vl_bvp.cuh")
print(f"{'other files':24s} inst={ii / ti * 100:5.1f}% samp={ss / ts * 100:5.1f}%")
for (f, ln, s), (ins, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{f}:{ln:4d} inst={ins / ti * 100:5.1f}% samp={sm / ts * 100:5.1f}%  {s.strip()[:100]}")
