#!/bin/bash
# Round-2 (second half) evidence run: smoke, bench line, sequence-length sweep, op tables (all ops, decode GEMMs, decode core,
# state update), launch list, full ncu captures of the forward scan kernel, the weight-streaming GEMM and the decode core.
#   usage: bash scripts/gpu_final_r2b.sh <tag>
tag=${1:-r2b}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
rm -f gpurun_out/sweep_$tag.jsonl
for L in 1024 4096 16384 65536; do
  timeout 120 python bench.py --no-cpu --workloads none --seqlen $L 2>/dev/null | tail -1 >> gpurun_out/sweep_$tag.jsonl
done
echo "sweep lines: $(wc -l < gpurun_out/sweep_$tag.jsonl)"
timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops_$tag.json 2> gpurun_out/bench_ops_$tag.err; echo "ops rc=$?"
timeout 100 python scripts/bench_skinny.py > gpurun_out/bench_skinny_$tag.json 2> gpurun_out/bench_skinny_$tag.err; echo "skinny rc=$?"
(timeout 100 python scripts/bench_decode_core.py; timeout 100 python scripts/bench_decode_core.py fp32; timeout 100 python scripts/bench_ssu.py) 2>/dev/null > gpurun_out/bench_decode_ops_$tag.jsonl; echo "decode ops rc=$?"
timeout 100 python scripts/trace_tc.py > gpurun_out/trace_fwd_$tag.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --workloads none --sustained-s 0 > gpurun_out/bench_ncu_$tag.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ssd_tc_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_fwd_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-bwd --workloads none --sustained-s 0 > gpurun_out/ncu_fwd_$tag.log 2>&1; echo "ncu fwd rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 20 -c 2 -f -o gpurun_out/prof_skinny_$tag \
    python scripts/prof_decode.py bf16 > gpurun_out/ncu_skinny_$tag.log 2>&1; echo "ncu skinny rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mamba2_decode_core -s 30 -c 1 -f -o gpurun_out/prof_dec_$tag \
    python scripts/prof_decode.py bf16 > gpurun_out/ncu_dec_$tag.log 2>&1; echo "ncu decode rc=$?"
ls -la gpurun_out/*_$tag* | awk '{print $5, $9}'
