mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 1200 gpurun_out/r02_bench_n8.json; grep -v "^W0\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02_bench_n8.err | tail -5
