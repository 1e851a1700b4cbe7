#!/bin/bash
# Quick experiment visit: tensor-core parity tests, fwd(+bwd) bench without the CPU leg, per-role trace.
tag=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc_$tag.txt
tail -3 gpurun_out/pytest_$tag.log
timeout 300 python bench.py --no-cpu --workloads none > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?" >> gpurun_out/rc_$tag.txt
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("fwd ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "fwd_bwd ms", d["fwd_bwd"]["ms_per_step"] if d.get("fwd_bwd") else None)
PY
timeout 120 python scripts/kernel_times.py > gpurun_out/ktimes_$tag.txt 2>&1; cat gpurun_out/ktimes_$tag.txt
timeout 120 python scripts/trace_tc.py > gpurun_out/trace_$tag.txt 2>&1; echo "trace rc=$?" >> gpurun_out/rc_$tag.txt
tail -1 gpurun_out/trace_$tag.txt
cat gpurun_out/rc_$tag.txt
