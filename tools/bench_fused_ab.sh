#!/bin/bash
# Development aid (GPU box): the latency-bound workload (BASELINE configs[0]) with variants of the library built by
# tools/build_variant.sh (e.g. nofused -DMPB_FUSED_DEFAULT=0: the stream-ordered loop).  usage: bench_fused_ab.sh tags...
cd "$(dirname "$0")/.."
for tag in "$@"; do
lib=build/libmpb200_$tag.so; [ "$tag" == "default" ] && lib=matching-pursuit_b200/libmpb200.so
MPB200_LIBRARY=$PWD/$lib python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; l=json.loads(sys.stdin.readline()); its=l['config']['iterations']
print('$tag', 'atoms/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'us_per_iteration', round(1e3*l['ms_per_step']/its,2), 'launches', l['gpu_launches'], 'loop_ms', round(l['kernel_ms']['recorrelate_per_iteration'],4), 'first_pass_ms', round(l['kernel_ms']['first_pass'],4))"
done
