mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generate_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 300 -k "generate or ask or continuation or caches or facade or load_pretrained" > gpurun_out/r02_c4_tests.log 2>&1; tail -25 gpurun_out/r02_c4_tests.log
L=gpurun_out/r02_c4_tower_long.log; : > $L
timeout 200 python tools/bench_tower.py --tower vit --reps 16 --tag default >> $L 2>&1
VIDI_GEMM2_RELAXED=0 timeout 200 python tools/bench_tower.py --tower vit --reps 16 --tag release_arrive >> $L 2>&1
timeout 200 python tools/bench_tower.py --tower vit --reps 16 --bn 256 --tag bn256 >> $L 2>&1
timeout 200 python tools/bench_tower.py --tower vit --reps 16 --bn 128 --tag bn128 >> $L 2>&1
timeout 200 python tools/bench_tower.py --tower vit --reps 32 --frames 64 --tag frames64 >> $L 2>&1
timeout 200 python tools/bench_tower.py --tower vit --reps 16 --no-attn --tag noattn >> $L 2>&1
timeout 200 python tools/bench_tower.py --tower aud --reps 40 --tag aud_default >> $L 2>&1
cat $L
