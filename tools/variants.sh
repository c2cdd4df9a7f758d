#!/bin/bash
# runs bench.py over every experimental build in variants/ (csrc/build.py --out=variants/libtr_X.so -D...) on the GPU box
for so in variants/libtr_*.so; do
  v=$(basename $so .so); v=${v#libtr_}
  TR_LIB=$PWD/$so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --min-seconds 0 "$@" > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python tools/benchline.py gpurun_out/bench_var_$v.json
done
