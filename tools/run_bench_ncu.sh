#!/bin/bash
# development helper, run under gpurun, ONE profiler pass per call:  tools/grun.sh 1200 'tools/run_bench_ncu.sh <what>'
#   bench     plain bench (both arms)
#   launches  ncu launch list (gpu__time_duration per launch) of a short bench run
#   gemm | bm25 | scan   ncu --set full capture of that kernel (second launch)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --verify 0 --docs ${NCU_DOCS:-10000000}"
case "$1" in
  bench)
    python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
    python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json ;;
  launches)
    $CMD > gpurun_out/plain.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
    echo "launch list rc=$?" ;;
  gemm)
    $CMD > gpurun_out/plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:dense_gemm_topk -s 1 -c 1 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu2.log 2>&1; echo "gemm prof rc=$?" ;;
  bm25)  # the 16-bit fast kernel of the second step (launches alternate <16>, <32>: skip two)
    $CMD > gpurun_out/plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:bm25_fast -s 2 -c 1 -f -o gpurun_out/prof_bm25 $CMD > gpurun_out/ncu3.log 2>&1; echo "bm25 prof rc=$?" ;;
  scan)
    timeout 300 python tools/gpu_probe.py perf_scan > gpurun_out/plain_scan.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:dense_scan_tma -s 1 -c 1 -f -o gpurun_out/prof_scan python tools/gpu_probe.py perf_scan > gpurun_out/ncu4.log 2>&1; echo "scan prof rc=$?"
    cat gpurun_out/plain_scan.log ;;
  *) echo "usage: $0 bench|launches|gemm|bm25|scan"; exit 2 ;;
esac
ls -la gpurun_out/ | tail -12
