# 8-GPU safety run: N-vs-1 parity at world 8 (IPC arenas with 8 peers), then the bench line at N=8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q --timeout 600 -k "8-" > gpurun_out/r02_dist_tests_n8.log 2>&1; tail -6 gpurun_out/r02_dist_tests_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 2200 gpurun_out/r02_bench_n8.json; grep -v "^W0\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02_bench_n8.err | tail -8
