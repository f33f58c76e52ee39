mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 200 -k "resize or premerge or peer_exchange or merge2" > gpurun_out/r02_c2_tests_a.log 2>&1; tail -12 gpurun_out/r02_c2_tests_a.log
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 300 -k "multirank or fake_multirank or mini_config1" > gpurun_out/r02_c2_tests_b.log 2>&1; tail -12 gpurun_out/r02_c2_tests_b.log
for r in 0 1; do VIDI_GEMM2_RELAXED=$r timeout 120 python tools/bench_tower.py --tower vit --tag relaxed$r; done 2>&1 | tee gpurun_out/r02_c2_tower_ab.log
VIDI_GEMM2_RELAXED=1 timeout 120 python tools/bench_tower.py --tower vit --tag relaxed1_noattn --no-attn 2>&1 | tee -a gpurun_out/r02_c2_tower_ab.log
VIDI_GEMM2_RELAXED=0 timeout 120 python tools/bench_tower.py --tower vit --tag 1cta --cta2 0 2>&1 | tee -a gpurun_out/r02_c2_tower_ab.log
for r in 0 1; do VIDI_GEMM2_RELAXED=$r timeout 120 python tools/bench_tower.py --tower aud --tag relaxed$r; done 2>&1 | tee -a gpurun_out/r02_c2_tower_ab.log
VIDI_GEMM2_RELAXED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 2 -c 1 -o gpurun_out/r02_prof_gemm2_vit_qkv -f python tools/bench_kernels.py one vit_qkv > gpurun_out/r02_prof_gemm2_vit_qkv.stdout 2>&1
ls -la gpurun_out | tail -5
