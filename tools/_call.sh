mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x ) > gpurun_out/c18_tests.log 2>&1
grep -E "passed|failed" gpurun_out/c18_tests.log | tail -2
timeout 300 python tools/bench_k12_large.py 2>&1 | head -1
ncu --set full --clock-control none --import-source on -k regex:"structure_factor_box_kernel|sf_box|sf_phase" -s 5 -c 5 -o gpurun_out/c18_k1 python tools/prof_dense.py large 512 > gpurun_out/c18_ncu.log 2>&1
tail -1 gpurun_out/c18_ncu.log
