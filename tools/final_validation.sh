#!/bin/bash
# Round-end validation on the GPU box: smoke, bench (both arms), ncu launch list + one full GEMM capture.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 2> gpurun_out/bench_r01.err | tail -1 > gpurun_out/bench_r01.json; cut -c1-400 gpurun_out/bench_r01.json
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --gpus 1 --steps 1 --warmup 1 2> gpurun_out/bench_ref_r01.err | tail -1 > gpurun_out/bench_ref_r01.json; cut -c1-600 gpurun_out/bench_ref_r01.json
echo "== ncu launch list (eager launches of the same bench command)"
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv \
  --log-file gpurun_out/launches_r01_final.csv python bench.py --gpus 1 --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu --set full, one FFN1-shaped GEMM launch"
B200ST_NO_PDL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 60 -c 1 -f -o gpurun_out/prof_gemm_r01_final \
  python tools/one_step.py 1 > gpurun_out/ncu_gemm_final.log 2>&1
tail -2 gpurun_out/ncu_gemm_final.log | cut -c1-200
