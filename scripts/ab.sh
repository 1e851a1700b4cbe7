#!/bin/bash
# Same-box A/B of library builds: scripts/ab.sh <tag> libA.so libB.so ...   (paths relative to the repo root; "cur" = the in-tree lib)
tag=$1; shift
mkdir -p gpurun_out
for round in 1 2; do
  for lib in "$@"; do
    if [ "$lib" = "cur" ]; then unset OMNI_LIB_PATH; else export OMNI_LIB_PATH=$PWD/$lib; fi
    echo "== $lib (round $round)"
    timeout 200 python scripts/kernel_times.py 2>/dev/null | head -3
    timeout 200 python bench.py --no-cpu --no-bwd --workloads none --sustained-s 0.5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bench fwd ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'sustained', d['roofline'].get('sustained',{}).get('ms_per_step'))"
  done
done 2>&1 | tee gpurun_out/ab_$tag.txt
