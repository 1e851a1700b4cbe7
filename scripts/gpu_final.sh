#!/bin/bash
# Round-end visit: smoke, all GPU parity tests, bench line (with the CPU leg), per-role trace, launch list and one full ncu
# capture of the forward scan kernel.   usage: bash scripts/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" > gpurun_out/rc_$tag.txt
tail -1 gpurun_out/smoke_$tag.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_$tag.txt
tail -3 gpurun_out/pytest_$tag.log
timeout 300 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?" >> gpurun_out/rc_$tag.txt
cat gpurun_out/bench_$tag.json
timeout 120 python scripts/trace_tc.py > gpurun_out/trace_$tag.txt 2>&1; echo "trace rc=$?" >> gpurun_out/rc_$tag.txt
tail -1 gpurun_out/trace_$tag.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ssd_tc_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_fwd_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-bwd > gpurun_out/ncu_fwd_$tag.log 2>&1
cat gpurun_out/rc_$tag.txt
