import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), {k:(round(v["ms"],4),round(v["fp32_frac"],3)) for k,v in d.get("kernels").items()}, "vis", round(d["passes_ms"]["visibility"],3))
    except Exception as e:
        print(f, "ERR", e)
