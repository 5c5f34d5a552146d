set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/c6_tests.log 2>&1
tail -4 gpurun_out/c6_tests.log
timeout 300 python tools/bench_k12_large.py > gpurun_out/c6_k12.txt 2>&1
cat gpurun_out/c6_k12.txt
timeout 300 python tools/bench_k3_variants.py 16384 > gpurun_out/c6_k3_variants.txt 2>&1
cat gpurun_out/c6_k3_variants.txt
