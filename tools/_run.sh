TRR_GEMM_PAIR=1 timeout 120 python tools/gpu_probe.py gemm_debug 2>&1 | tail -8
timeout 300 python tools/gpu_probe.py gemm_path 2>&1 | tail -5
python -m pytest tests/test_gpu_dense.py tests/test_gpu_hybrid.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/gpu_probe.py perf_gemm 2>&1 | tail -3
TRR_GEMM_PAIR=0 timeout 300 python tools/gpu_probe.py perf_gemm 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kernels')}); print(d['roofline']['frac'])"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --docs 4000000"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bm25_search_fast -c 1 -o gpurun_out/prof_bm25fast2 $CMD > gpurun_out/ncu_b2.log 2>&1; echo ncu rc=$?
