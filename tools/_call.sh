mkdir -p gpurun_out
DIFFSIMS_B200_LIB=tools/microbench/_libprof.so timeout 300 python tools/prof_rows_roles.py > gpurun_out/r02_rows_roles.txt 2>&1
DIFFSIMS_B200_LIB=tools/microbench/_libprof.so timeout 300 python tools/prof_umma_roles.py > gpurun_out/r02_umma_roles_v8.txt 2>&1
tail -4 gpurun_out/r02_rows_roles.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err
tail -c 300 gpurun_out/r02c_bench_2gpu.json; tail -3 gpurun_out/r02c_bench_2gpu.err
