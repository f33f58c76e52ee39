mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generate_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 300 -k "native or vidi7b or multirank or facade" > gpurun_out/r02_c5_tests.log 2>&1; tail -15 gpurun_out/r02_c5_tests.log
L=gpurun_out/r02_c5_bench_ab.log; : > $L
for v in "" "--ln-fold" "--attn-poly 4" "--attn-poly 3"; do
  echo "== bench --quick --steps 3 $v" >> $L
  timeout 300 python bench.py --quick --steps 3 --no-cpu-baseline $v >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c5_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    print(d["ms_per_step"], d["value"], "tower", d["roofline"]["by_site"].get("tower"), "frac", d["roofline"]["frac"], {k:v for k,v in d["other_ops_ms_per_step"].items() if v>5}, "text", d.get("text_pass_ms"), d["clocks"])
PY
