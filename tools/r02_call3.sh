# 2-GPU call: torchrun N-vs-1 parity (both exchange modes), then the bench at N=2 (parity pre-flight + digest) and at N=1 for the digest
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r02_c3_dist_tests.log 2>&1; tail -25 gpurun_out/r02_c3_dist_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c3_bench_n2.json 2> gpurun_out/r02_c3_bench_n2.err; tail -c 3000 gpurun_out/r02_c3_bench_n2.json; tail -5 gpurun_out/r02_c3_bench_n2.err
