mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_generate_gpu.py -m gpu -q --timeout 400 > gpurun_out/r02_c14_tests.log 2>&1; tail -8 gpurun_out/r02_c14_tests.log
L=gpurun_out/r02_c14_bench_ab.log; : > $L
for v in "VIDI_GEMM_TMASTORE=1 VIDI_LN_ROWS=1" "VIDI_GEMM_TMASTORE=0 VIDI_LN_ROWS=1" "VIDI_GEMM_TMASTORE=1 VIDI_LN_ROWS=0" "VIDI_GEMM_TMASTORE=1 VIDI_LN_ROWS=1"; do
  echo "== $v bench --quick --steps 3" >> $L
  env $v timeout 300 python bench.py --quick --steps 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c14_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    bs=d["roofline"]["by_site"]
    print(d["ms_per_step"], d["value"], {k:(v["tflops"],v["ms_per_step"]) for k,v in bs.items() if k.startswith("llm") or k=="tower"}, "frac", d["roofline"]["frac"], {k:v for k,v in d["other_ops_ms_per_step"].items() if v>50}, d["clocks"]["sm_mhz"], d["clocks"]["avg_power_w"], d["logits_digest"]["top5_logits"])
PY
