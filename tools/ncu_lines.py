"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report captured with --import-source on:
python tools/ncu_lines.py report.ncu-rep kernel_substring [min_pct]"""
import csv, subprocess, sys, io
rep, kn = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
funcs = []; cur = None; fpath = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fpath = r[1]
    elif len(r) >= 2 and r[0] == "Function Name":
        cur = {"name": r[1], "rows": [], "hdr": None, "file": fpath}; funcs.append(cur)
    elif cur is not None and r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]): cur["rows"].append(r)
L = {}
for f in funcs:
    if kn not in f["name"]: continue
    h = f["hdr"]; iI = h.index("Instructions Executed"); iS = h.index("# Samples")
    for r in f["rows"]:
        if r[0].isdigit():
            k = (f["file"].split('/')[-1], int(r[0]))
            o = L.setdefault(k, [0, 0, r[1]]); o[0] += int(r[iI] or 0); o[1] += int(r[iS] or 0)
T = sum(v[0] for v in L.values()); S = sum(v[1] for v in L.values())
print(f"{kn}: {T} warp instructions, {S} samples")
for k in sorted(L):
    v = L[k]
    if v[0] / T * 100 >= minpct or v[1] / S * 100 >= minpct:
        print(f"{k[0][:18]:18s} {k[1]:5d} inst {v[0]/T*100:5.2f}% smp {v[1]/S*100:5.2f}%  {v[2][:100]}")
