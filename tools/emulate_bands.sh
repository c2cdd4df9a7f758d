#!/bin/bash
# one GPU renders each band of an N-way split in turn (bench.py --emulate-band r/N): the per-pass device times a rank would
# see, without paying for an N-GPU box
N=${1:-8}; shift
for r in $(seq 0 $((N-1))); do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 0 --emulate-band $r/$N "$@" > gpurun_out/band_${r}_of_$N.json 2> gpurun_out/band_${r}_of_$N.err
  python - <<P
import json
d=json.loads(open("gpurun_out/band_${r}_of_$N.json").read().strip().splitlines()[-1])
print("band $r/$N", d.get("band_rows"), round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["passes_ms"].items()})
P
done
