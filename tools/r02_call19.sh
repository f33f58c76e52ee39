mkdir -p gpurun_out
L=gpurun_out/r02_c19_bench_ab.log; : > $L
for v in 48 24 32 16; do
  echo "== VIDI_GEMM2_L2MB=$v bench --quick --steps 3" >> $L
  VIDI_GEMM2_L2MB=$v timeout 300 python bench.py --quick --steps 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_c19_bench_ab.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:300]); continue
    d=json.loads(line)
    bs=d["roofline"]["by_site"]
    print(d["ms_per_step"], d["value"], {k:(v["tflops"],v["ms_per_step"]) for k,v in bs.items() if k.startswith("llm") or k=="tower"}, "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["avg_power_w"])
PY
VIDI_GEMM2_L2MB=24 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm2_bf16_kernel -s 2 -c 1 python tools/bench_kernels.py one gate_up126k_2cta 2>&1 | grep -E "dram__|duration" 
VIDI_GEMM2_L2MB=16 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm2_bf16_kernel -s 2 -c 1 python tools/bench_kernels.py one gate_up126k_2cta 2>&1 | grep -E "dram__|duration"
