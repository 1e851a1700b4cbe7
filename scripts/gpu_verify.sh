#!/bin/bash
# Round-end verification on one B200, as the driver runs it: full GPU test suite, smoke(), default bench line, reference arm.
#   usage: bash scripts/gpu_verify.sh <tag>
tag=${1:-v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_full_$tag.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("fwd ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "fwd_bwd", d["fwd_bwd"]["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d.get("workloads", {}).items(): print(k, v.get("ms_per_step"))
print(open("gpurun_out/bench_ref_$tag.json").read()[:400])
PY
