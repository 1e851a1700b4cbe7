#!/bin/bash
# ncu --set full of one launch of each streaming kernel (conv1d fwd, gated norm fwd, state update) through scripts/bench_ops.py
mkdir -p gpurun_out
for k in conv1d_fwd_kernel norm_gated_fwd_kernel ssu_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python scripts/bench_ops.py > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
