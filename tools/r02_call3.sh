# multi-GPU call: torchrun N-vs-1 parity (both exchange modes, mini + C3-shaped), then the bench at N (parity pre-flight + digest)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q --timeout 600 > gpurun_out/r02_dist_tests_n$N.log 2>&1; tail -12 gpurun_out/r02_dist_tests_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 2500 gpurun_out/r02_bench_n$N.json; grep -v "^W0\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02_bench_n$N.err | tail -8
