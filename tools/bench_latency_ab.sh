#!/bin/bash
cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
for lib in build/libmpb200_nopdl.so matching-pursuit_b200/libmpb200.so; do for w in c1 c2; do
MPB200_LIBRARY=$PWD/$lib python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; l=json.loads(sys.stdin.readline()); its=l['config']['iterations']
print('$lib', '$w', 'atoms/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'us_per_iteration', round(1e3*l['ms_per_step']/its,2), l['kernel_ms'], l['config']['mode'])"
done; done
