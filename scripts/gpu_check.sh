#!/bin/bash
# One GPU-box visit: parity tests, bench line, launch list, per-role trace of the forward kernel.
# usage (from the repo root on the box): bash scripts/gpu_check.sh <tag> [full]
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc_$tag.txt
tail -3 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?" >> gpurun_out/rc_$tag.txt
cat gpurun_out/bench_$tag.json
python scripts/trace_tc.py > gpurun_out/trace_$tag.txt 2>&1; echo "trace rc=$?" >> gpurun_out/rc_$tag.txt
tail -2 gpurun_out/trace_$tag.txt
if [ "$2" = "full" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu_$tag.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ssd_tc_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_fwd_$tag \
      python bench.py --steps 1 --warmup 3 --no-cpu --no-bwd > gpurun_out/ncu_fwd_$tag.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ssd_tc_bwd -s 1 -c 1 -f -o gpurun_out/prof_bwd_$tag \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bwd_$tag.log 2>&1
fi
cat gpurun_out/rc_$tag.txt
