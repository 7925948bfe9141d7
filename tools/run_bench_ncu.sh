#!/bin/bash
# plain bench (both arms), then the ncu launch list and full captures of the hot kernels (development helper;
# run under gpurun: tools/grun.sh 2400 'tools/run_bench_ncu.sh')
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 --docs ${NCU_DOCS:-10000000}"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dense_gemm_topk -s 1 -c 1 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu2.log 2>&1; echo "gemm prof rc=$?"
ncu --set full --clock-control none --import-source on -k regex:bm25_search -s 1 -c 1 -o gpurun_out/prof_bm25 $CMD > gpurun_out/ncu3.log 2>&1; echo "bm25 prof rc=$?"
timeout 300 python tools/gpu_probe.py perf_scan > gpurun_out/plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dense_scan_tma -s 1 -c 1 -o gpurun_out/prof_scan python tools/gpu_probe.py perf_scan > gpurun_out/ncu4.log 2>&1; echo "scan prof rc=$?"
cat gpurun_out/plain_scan.log
ls -la gpurun_out/
