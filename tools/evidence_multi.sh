#!/bin/bash
# Multi-GPU evidence on one N-GPU box (gpurun --gpus 8 -- tools/evidence_multi.sh r02): the 4K frame band-parallel at N = 8 / 4 / 2 and
# BASELINE.json configs[4] (64 views at 8K) split views x bands.  Results in gpurun_out/<tag>/; every run is bounded by its own timeout.
tag=${1:-r02}; out=gpurun_out/$tag; mkdir -p $out
G=$(nvidia-smi -L | wc -l)
run() {  # run N name args...
  local n=$1 name=$2; shift 2
  [ $n -le $G ] || return 0
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $n "$@" > $out/${tag}_bench_$name.json 2> $out/$name.err
  echo "$name rc=$?"; python tools/benchline.py $out/${tag}_bench_$name.json | cut -c1-150
}
run 8 4k_n8
run 8 config5_views1x_bands8 --workload config5 --view-groups 1 --steps 2 --warmup 3 --min-seconds 0
run 8 config5_views8x_bands1 --workload config5 --view-groups 8 --steps 2 --warmup 3 --min-seconds 0
run 4 4k_n4
run 8 config5_views2x_bands4 --workload config5 --view-groups 2 --steps 2 --warmup 3 --min-seconds 0
run 8 config5_views4x_bands2 --workload config5 --view-groups 4 --steps 2 --warmup 3 --min-seconds 0
run 2 4k_n2
