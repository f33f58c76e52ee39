# every BASELINE workload on one GPU (quick lines: 1 warm-up + 2 timed steps, no e2e / cpu legs) -> table in DESIGN.md
mkdir -p gpurun_out
L=gpurun_out/r02_workloads_n1.log; : > $L
for w in c1 c2 c2p c5 c4; do
  echo "== $w" >> $L
  timeout 400 python bench.py --workload $w --quick --steps 2 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_workloads_n1.log"):
    if line.startswith("=="): print(line.strip()); continue
    if not line.startswith("{"): print(line.strip()[:200]); continue
    d=json.loads(line)
    print(d["config"]["workload"][:60], "| tokens", d["config"]["total_tokens"], "| ms/step", d["ms_per_step"], "| tok/s", d["value"], "| frac", d["roofline"]["frac"], "| clocks", d["clocks"]["sm_mhz"], d["clocks"].get("avg_power_w"))
PY
