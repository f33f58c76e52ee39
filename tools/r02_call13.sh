mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 400 -x -k "c3_shape or pos_table" > gpurun_out/r02_c13_tests_a.log 2>&1; tail -5 gpurun_out/r02_c13_tests_a.log
VIDI_GEMM2_TMASTORE=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 300 -x -k "gemm or prefill_mini or true_dims or ln_fold or vidi7b" > gpurun_out/r02_c13_tests_ts.log 2>&1; tail -8 gpurun_out/r02_c13_tests_ts.log
L=gpurun_out/r02_c13_tower_ts.log; : > $L
for r in 0 1 0 1; do VIDI_GEMM2_TMASTORE=$r timeout 200 python tools/bench_tower.py --tower vit --reps 16 --tag tmastore$r >> $L 2>&1; done
VIDI_GEMM2_TMASTORE=1 timeout 200 python tools/bench_tower.py --tower aud --reps 40 --tag aud_tmastore1 >> $L 2>&1
VIDI_GEMM2_TMASTORE=0 timeout 200 python tools/bench_tower.py --tower aud --reps 40 --tag aud_tmastore0 >> $L 2>&1
cat $L
L=gpurun_out/r02_c13_bench_ab.log; : > $L
for r in 1 0; do
  echo "== VIDI_GEMM2_TMASTORE=$r bench --quick --steps 3" >> $L
  VIDI_GEMM2_TMASTORE=$r timeout 300 python bench.py --quick --steps 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c13_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    bs=d["roofline"]["by_site"]
    print(d["ms_per_step"], d["value"], {k:(v["tflops"],v["ms_per_step"]) for k,v in bs.items() if k=="tower"}, "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["clocks"], d["logits_digest"]["top5_logits"])
PY
