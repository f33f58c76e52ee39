mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_generate_gpu.py -m gpu -q --timeout 300 -x -k "gemm or native or facade or generate or prefill_mini or vidi7b or continuation" > gpurun_out/r02_c8_tests.log 2>&1; tail -12 gpurun_out/r02_c8_tests.log
timeout 300 python tools/bench_kernels.py text 2>&1 | tee gpurun_out/r02_c8_text_gemm.log
L=gpurun_out/r02_c8_bench_ab.log; : > $L
for r in 1 0; do
  echo "== VIDI_GEMM_SKINNY=$r bench --steps 3 --warmup 3" >> $L
  VIDI_GEMM_SKINNY=$r timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c8_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "text_pass_ms", d["text_pass_ms"], "decode", d["decode"], "text gemm", d["roofline"]["by_site"].get("text"), {k:v for k,v in d["other_ops_ms_per_step"].items() if "xattn" in k or "text" in k})
PY
