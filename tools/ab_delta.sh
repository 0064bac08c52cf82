#!/bin/bash
# Development aid (GPU box): A/B timing of libmpb200 variants built by tools/build_variant.sh on a slice of the
# headline workload.  usage: tools/ab_delta.sh [batch] [iterations] -- tags...   (tag "default" = the in-tree library)
cd "$(dirname "$0")/.."
batch=${1:-128}; iters=${2:-32}; shift 2; [ "$1" == "--" ] && shift
mkdir -p gpurun_out
for tag in "$@"; do
  lib=build/libmpb200_$tag.so; [ "$tag" == "default" ] && lib=matching-pursuit_b200/libmpb200.so
  for rep in 1 2; do
    MPB200_LIBRARY=$PWD/$lib python bench.py --batch $batch --iterations $iters --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
      2>>gpurun_out/ab_err.log | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('$tag', 'batch', $batch, 'rep', $rep, 'delta_ms', round(l['kernel_ms']['gram_update_per_iteration'], 4), 'first_pass_ms', round(l['kernel_ms']['first_pass'], 2), 'apply_ms', round(l['kernel_ms']['apply_per_iteration'], 4), 'frac', round(l['roofline']['frac'], 4), 'sm_mhz', l['clocks']['sm_mhz'], 'atoms/s', round(l['value']))
"
  done
done
