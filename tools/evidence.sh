#!/bin/bash
# The round's measured evidence, one GPU: run on the GPU box (gpurun -- tools/evidence.sh r02), results in gpurun_out/<tag>/.
# Bench lines are taken first, never under a profiler; the ncu passes and the sanitizer come after.
tag=${1:-r02}; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_gpu_tests.txt
python bench.py > $out/${tag}_bench_4k_n1.json 2> $out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_4k_reference_arm.json 2>> $out/bench.err
for w in config1 config2 config3 8k; do
  python bench.py --workload $w --steps 30 --warmup 5 > $out/${tag}_bench_${w}.json 2>> $out/bench.err
done
python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline --min-seconds 0 > $out/${tag}_bench_config5_n1.json 2>> $out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches_bench_4k.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --min-seconds 0 > /dev/null 2>&1
for r in 1 5; do   # one rank's share of an 8-GPU frame: where the band path's time goes
  ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $out/${tag}_launches_band_${r}of8.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --min-seconds 0 --emulate-band $r/8 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on \
    -k regex:"shade_kernel|raster_tiles|resolve_kernel|mip_kernel|tonemap_kernel|bin_count|bin_fill|cull_kernel|assign_lights" -s 12 -c 11 \
    -o $out/${tag}_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --min-seconds 0 > $out/ncu_full.log 2>&1
for t in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $t --error-exitcode 7 python __graft_entry__.py --smoke > $out/${tag}_${t}_smoke.log 2>&1
  echo "$t rc=$?" >> $out/sanitizer_rc.txt
done
cat $out/${tag}_gpu_tests.txt $out/sanitizer_rc.txt
for f in $out/${tag}_bench_4k_n1.json $out/${tag}_bench_config*.json $out/${tag}_bench_8k.json; do python tools/benchline.py $f | cut -c1-260; done
