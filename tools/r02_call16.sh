mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 400 -x -k "gemm or prefill or vidi7b" > gpurun_out/r02_c16_tests.log 2>&1; tail -5 gpurun_out/r02_c16_tests.log
L=gpurun_out/r02_c16_bench_ab.log; : > $L
for v in "--llm-cta2 1" "--llm-cta2 0" "--llm-cta2 1"; do
  echo "== bench --quick --steps 3 $v" >> $L
  timeout 300 python bench.py --quick --steps 3 --no-cpu-baseline $v >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c16_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    bs=d["roofline"]["by_site"]
    print(d["ms_per_step"], d["value"], {k:(v["tflops"],v["ms_per_step"]) for k,v in bs.items() if k.startswith("llm") or k=="tower"}, "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["logits_digest"]["top5_logits"])
PY
