"""Copies a round's evidence from gpurun_out/<tag>/ (tools/evidence.sh) into profiles/ (tracked): bench lines, launch list,
ncu summary, DRAM traffic per launch (bench.py's roofline.traffic reads it), sanitizer logs, SASS mnemonic counts per kernel.
    python tools/collect_profiles.py r02"""
import csv, json, os, re, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
for f in sorted(os.listdir(src)):
    if f.startswith(tag + "_bench") or f.startswith(tag + "_launches") or f == tag + "_gpu_tests.txt":
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    if f.endswith("_smoke.log"):
        os.makedirs(os.path.join(dst, "sanitizer"), exist_ok=True)
        shutil.copy(os.path.join(src, f), os.path.join(dst, "sanitizer", f))
rep = os.path.join(src, tag + "_full.ncu-rep")
if os.path.exists(rep):
    out = os.path.join(dst, tag + "_ncu_full_summary.csv")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, out])
    rows = list(csv.DictReader(open(out)))
    names = {"shade_kernel<1": "shade_transmission", "shade_kernel<0": "shade_opaque", "mip_kernel": "mips", "tonemap_kernel": "tonemap",
             "raster_tiles": "raster_tiles", "resolve_kernel": "resolve", "bin_count": "bin_count", "bin_fill": "bin_fill"}
    traffic = {}
    for r in rows:
        for k, n in names.items():
            if r["kernel"].startswith(k) and n not in traffic:
                traffic[n] = (float(r["dram_rd_MB"]) + float(r["dram_wr_MB"])) * 1e6
    json.dump({"workload": "4k", "n_gpus": 1, "source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full, dram__bytes_read.sum + "
               "dram__bytes_write.sum per launch)", "bytes_per_launch": traffic}, open(os.path.join(dst, tag + "_dram_traffic.json"), "w"), indent=1)
# SASS evidence: per kernel, how many TMA bulk copies / mbarrier waits / packed-FP32 / FP64 instructions the shipped binary holds
so = os.path.join(ROOT, "transmission_renderer_b200", "libtr.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, counts = None, {}
pats = {"UBLKCP (TMA bulk copy)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "FFMA2/FMUL2/FADD2 (packed fp32)": r"\bF(FMA|MUL|ADD)2\b",
        "FFMA": r"\bFFMA\b", "MUFU": r"\bMUFU", "DFMA/DMUL/DADD (fp64)": r"\bD(FMA|MUL|ADD)\b", "REDUX": r"\bREDUX", "ATOMS/ATOMG/RED": r"\b(ATOMS|ATOMG|RED)\b",
        "LDS": r"\bLDS", "STS": r"\bSTS", "LDG": r"\bLDG", "STG": r"\bSTG", "BAR.SYNC": r"\bBAR\."}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", cur)[:90]
        counts[cur] = {k: 0 for k in pats}
    elif cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
with open(os.path.join(dst, tag + "_sass_mnemonics.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel (cuobjdump -sass libtr.so, sm_100a)"] + list(pats))
    for k, v in counts.items():
        w.writerow([k] + [v[p] for p in pats])
print("profiles/ updated:", sorted(x for x in os.listdir(dst) if x.startswith(tag)))
