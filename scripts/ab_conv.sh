#!/bin/bash
# Same-box A/B of conv1d-backward builds: scripts/ab_conv.sh <tag> name1 name2 ...   (names of .ab/lib_<name>.so; "cur" = in-tree)
tag=$1; shift
mkdir -p gpurun_out
{
for round in ${ROUNDS:-1 2}; do
  for name in "$@"; do
    if [ "$name" = "cur" ]; then unset OMNI_LIB_PATH; else export OMNI_LIB_PATH=$PWD/.ab/lib_$name.so; fi
    echo "== $name (round $round)"
    timeout 120 python scripts/bench_conv_bwd.py 2>&1 | tail -6
  done
done
unset OMNI_LIB_PATH
} 2>&1 | tee gpurun_out/ab_conv_$tag.txt
