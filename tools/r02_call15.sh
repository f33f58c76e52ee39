mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 400 -x -k "gemm or prefill or vidi7b or ln_fold or multirank" > gpurun_out/r02_c15_tests.log 2>&1; tail -6 gpurun_out/r02_c15_tests.log
L=gpurun_out/r02_c15_tower_rt.log; : > $L
for r in 1 0 1 0; do VIDI_GEMM2_RESTMA=$r timeout 200 python tools/bench_tower.py --tower vit --reps 16 --tag restma$r >> $L 2>&1; done
cat $L
L=gpurun_out/r02_c15_bench_ab.log; : > $L
for r in 1 0; do
  echo "== VIDI_GEMM2_RESTMA=$r bench --quick --steps 3" >> $L
  VIDI_GEMM2_RESTMA=$r timeout 300 python bench.py --quick --steps 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c15_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    bs=d["roofline"]["by_site"]
    print(d["ms_per_step"], d["value"], {k:(v["tflops"],v["ms_per_step"]) for k,v in bs.items() if k=="tower"}, "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["avg_power_w"], d["logits_digest"]["top5_logits"])
PY
